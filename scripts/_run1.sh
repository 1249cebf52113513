set -x
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -x -k attention 2>&1 | tail -5
SRW_ATTN_FWD=smem SRW_ATTN_BWD=smem timeout 100 python scripts/attn_bench.py 2>&1 | tail -2
timeout 100 python scripts/attn_bench.py 2>&1 | tail -2
timeout 100 python scripts/attn_bench.py --B 16 2>&1 | tail -2
timeout 120 python scripts/attn_trace.py > gpurun_out/attn_trace_v4.txt 2>&1; cat gpurun_out/attn_trace_v4.txt
