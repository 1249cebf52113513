"""Microbenchmark of srw_gemm on the ViT-S step's shapes (CUDA events, L2-cold rotation over several buffers).
python scripts/gemm_bench.py [--reps 20] [--only NAME]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semireward_b200 import _lib as L, ops as O  # noqa: E402

T, TG, D, F = 6168, 4112, 384, 1536
SHAPES = {  # name: (M, N, K, a_mn, b_mn, epilogue, split_k)
    "qkv_fwd": (T, 3 * D, D, 0, 0, L.EPI_PLANES, 1),
    "proj_fwd": (T, D, D, 0, 0, L.EPI_RESID, 1),
    "fc1_fwd": (T, F, D, 0, 0, L.EPI_GELU, 1),
    "fc2_fwd": (T, D, F, 0, 0, L.EPI_RESID, 1),
    "fc2_dgrad": (TG, F, D, 0, 1, L.EPI_DGELU, 1),
    "fc1_dgrad": (TG, D, F, 0, 1, L.EPI_F32, 1),
    "qkv_dgrad": (TG, D, 3 * D, 0, 1, L.EPI_F32, 1),
    "fc2_wgrad": (D, F, TG, 1, 1, L.EPI_SPLITK, 8),
    "fc1_wgrad": (F, D, TG, 1, 1, L.EPI_SPLITK, 8),
    "qkv_wgrad": (3 * D, D, TG, 1, 1, L.EPI_SPLITK, 10),
    "proj_wgrad": (D, D, TG, 1, 1, L.EPI_SPLITK, 16),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default=None)
    ap.add_argument("--epi", type=int, default=None, help="override the epilogue (0 F32, 1 PLANES, 2 GELU, 3 RESID, 4 DGELU)")
    ap.add_argument("--impl", type=int, default=0)
    a = ap.parse_args()
    torch.manual_seed(0)
    for name, (M, N, K, amn, bmn, epi, split) in SHAPES.items():
        if a.only and a.only != name:
            continue
        if a.epi is not None and epi != L.EPI_SPLITK:
            epi = a.epi
        nbuf = 4
        As = [O.split_planes(torch.randn((K, M) if amn else (M, K), device="cuda")) for _ in range(nbuf)]
        Bs = [O.split_planes(torch.randn((K, N) if bmn else (N, K), device="cuda") * 0.05) for _ in range(nbuf)]
        bias = torch.randn(N, device="cuda")
        resid = torch.randn(M, N, device="cuda") if epi == L.EPI_RESID else None
        aux = torch.randn(M, N, device="cuda") if epi == L.EPI_DGELU else None
        outf = torch.empty(M, N, device="cuda") if epi in (L.EPI_F32, L.EPI_GELU, L.EPI_RESID) else None
        outp = O.empty_planes(M, N) if epi in (L.EPI_PLANES, L.EPI_GELU, L.EPI_DGELU) else None
        ws = torch.empty(split, M, N, device="cuda") if epi == L.EPI_SPLITK else None

        def run(i):
            O.gemm(As[i % nbuf], Bs[i % nbuf], M, N, K, a_mn=bool(amn), b_mn=bool(bmn), epilogue=epi, bias=None if epi in (L.EPI_SPLITK, L.EPI_DGELU) else bias,
                   resid=resid, aux=aux, out_f32=outf, out_planes=outp, split_k=split, workspace=ws, impl=a.impl)
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.reps):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / a.reps
        tf = 2.0 * M * N * K / (us * 1e-6) / 1e12
        print(f"{name:12s} M={M:5d} N={N:5d} K={K:5d} {us:8.1f} us  {tf:7.1f} TFLOP/s algorithmic ({3 * tf:7.1f} issued)")


if __name__ == "__main__":
    main()
