#!/bin/bash
# Round-end validation on a GPU box (gpurun -- 'bash scripts/gpu_validate.sh'): parity tests, smoke, the bench line, the ncu
# launch list of the bench command and one `--set full` capture of the attention kernels.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 100 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; cut -c1-400 gpurun_out/final_bench.json; tail -2 gpurun_out/final_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-200
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 82 -c 6 -f -o gpurun_out/attn python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_attn.log 2>&1; tail -1 gpurun_out/ncu_attn.log | cut -c1-200
ls -la gpurun_out | tail -8
