#!/bin/bash
# Data-parallel knobs of the headline step at N GPUs (gpurun --gpus N -- 'bash scripts/dp_sweep.sh N'): overlap ranges of the backward
# (SRW_DP_SPLIT; 1 = one all-reduce after the backward), NCCL's CTA budget and algorithm.  One bench line per variant.
N=${1:-4}
run() { echo "== $*"; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 200 --warmup 10 --no-roofline 2>/dev/null | cut -c136-172; }
run SRW_DP_SPLIT=1
run SRW_DP_SPLIT=1 NCCL_MIN_CTAS=32
run SRW_DP_SPLIT=1 NCCL_MIN_CTAS=64
run SRW_DP_SPLIT=1 NCCL_ALGO=NVLS
run SRW_DP_SPLIT=1 NCCL_ALGO=Tree
run SRW_DP_SPLIT=2
run SRW_DP_SPLIT=2 NCCL_MIN_CTAS=32
