#!/bin/bash
# Round-end validation on a GPU box (gpurun -- 'bash scripts/gpu_validate.sh'): parity tests, smoke, the bench line, the ncu launch
# list of the bench command and `--set full` captures of the GEMM and attention kernels.  Outputs under gpurun_out/ (prefix $1).
P=${1:-final}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 100 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; cut -c1-400 gpurun_out/${P}_bench.json; tail -2 gpurun_out/${P}_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-eager-leg > gpurun_out/${P}_ncu_launch.log 2>&1; tail -1 gpurun_out/${P}_ncu_launch.log | cut -c1-200
# `--set full` captures: skip the set-up + warm-up steps, then the attention kernels of one step / the GEMMs of the first blocks of
# one forward and the last blocks of one backward.  The .ncu-rep files are summarised here (scripts/ncu_summary.py) and removed:
# gpurun copies back at most 64 MiB.
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 540 -c 36 -f -o /tmp/${P}_attn python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-eager-leg > gpurun_out/${P}_ncu_attn.log 2>&1; tail -1 gpurun_out/${P}_ncu_attn.log | cut -c1-200
python scripts/ncu_summary.py /tmp/${P}_attn.ncu-rep > gpurun_out/${P}_ncu_attn.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm -s 2190 -c 60 -f -o /tmp/${P}_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-eager-leg > gpurun_out/${P}_ncu_gemm.log 2>&1; tail -1 gpurun_out/${P}_ncu_gemm.log | cut -c1-200
python scripts/ncu_summary.py /tmp/${P}_gemm.ncu-rep > gpurun_out/${P}_ncu_gemm.txt 2>&1
ls -la gpurun_out | grep ${P}
