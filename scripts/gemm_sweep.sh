#!/bin/bash
# Per-shape sweep of the GEMM tile choice (SRW_GEMM_FORCE) on the ViT-S step's shapes: prints the default line and every forced variant.
for f in "" 1:64 1:128 1:192 2:128 2:192 2:256; do
  echo "== SRW_GEMM_FORCE=$f"
  SRW_GEMM_FORCE=$f python scripts/gemm_bench.py --reps 40 2>&1 | grep -v "^$"
done
