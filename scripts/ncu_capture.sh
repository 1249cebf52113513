#!/bin/bash
# ncu evidence of the headline step (gpurun -- 'bash scripts/ncu_capture.sh <prefix>'): launch list of the bench command and `--set full`
# captures of the attention kernels of one step and of the GEMMs of ~4 blocks, summarised on the box (the .ncu-rep files exceed what
# gpurun copies back).
P=${1:-r2_final}
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-eager-leg > gpurun_out/${P}_ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 540 -c 36 -f -o /tmp/${P}_attn python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-eager-leg > gpurun_out/${P}_ncu_attn.log 2>&1
python scripts/ncu_summary.py /tmp/${P}_attn.ncu-rep > gpurun_out/${P}_ncu_attn.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm -s 2190 -c 60 -f -o /tmp/${P}_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-eager-leg > gpurun_out/${P}_ncu_gemm.log 2>&1
python scripts/ncu_summary.py /tmp/${P}_gemm.ncu-rep > gpurun_out/${P}_ncu_gemm.txt 2>&1
tail -2 gpurun_out/${P}_ncu_gemm.txt
