set -x
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
SRW_FOLD=0 timeout 300 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/t30_nofold.json 2> gpurun_out/t30_nofold.err; cut -c1-220 gpurun_out/t30_nofold.json
timeout 300 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/t30_fold.json 2> gpurun_out/t30_fold.err; cut -c1-220 gpurun_out/t30_fold.json
tail -3 gpurun_out/t30_fold.err
