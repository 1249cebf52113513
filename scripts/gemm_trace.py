"""Per-CTA clock64 timeline of the 1-CTA tcgen05 GEMM (debug hook srw_gemm_set_trace) on the ViT-S step's shapes.
python scripts/gemm_trace.py [--only NAME]   (diagnostic; impl=2 forces the 1-CTA kernel)"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semireward_b200 import _lib as L, ops as O  # noqa: E402
from gemm_bench import SHAPES  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    lib = L.load()
    lib.srw_gemm_set_trace.argtypes = [C.c_void_p]
    lib.srw_gemm_set_trace.restype = C.c_int
    torch.manual_seed(0)
    ghz = 1.9
    for name, (M, N, K, amn, bmn, epi, split) in SHAPES.items():
        if a.only and a.only != name:
            continue
        A = O.split_planes(torch.randn((K, M) if amn else (M, K), device="cuda"))
        Bm = O.split_planes(torch.randn((K, N) if bmn else (N, K), device="cuda") * 0.05)
        bias = torch.randn(N, device="cuda")
        resid = torch.randn(M, N, device="cuda") if epi == L.EPI_RESID else None
        aux = torch.randn(M, N, device="cuda") if epi == L.EPI_DGELU else None
        outf = torch.empty(M, N, device="cuda") if epi in (L.EPI_F32, L.EPI_GELU, L.EPI_RESID) else None
        outp = O.empty_planes(M, N) if epi in (L.EPI_PLANES, L.EPI_GELU, L.EPI_DGELU) else None
        ws = torch.empty(split, M, N, device="cuda") if epi == L.EPI_SPLITK else None
        trace = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

        def run():
            O.gemm(A, Bm, M, N, K, a_mn=bool(amn), b_mn=bool(bmn), epilogue=epi, bias=None if epi in (L.EPI_SPLITK, L.EPI_DGELU) else bias,
                   resid=resid, aux=aux, out_f32=outf, out_planes=outp, split_k=split, workspace=ws, impl=2)
        for _ in range(3):
            run()
        for mode in ("warm-L2", "cold-L2"):
            if mode == "cold-L2":
                flush.zero_()
            torch.cuda.synchronize()
            lib.srw_gemm_set_trace(trace.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            lib.srw_gemm_set_trace(None)
            t = trace.cpu().numpy().reshape(148, 16).astype(np.float64)
            t = t[t[:, 15] > 0]
            rel = (t - t[:, :1]) / (ghz * 1e3)   # us since CTA entry
            med = lambda c: float(np.median(rel[:, c][t[:, c] > 0])) if (t[:, c] > 0).any() else float("nan")
            tiles = int(max(((t[:, 4:14:2] > 0).sum(axis=1))))
            line = (f"{name:11s} {mode}: kernel {e0.elapsed_time(e1) * 1e3:6.1f} us | CTAs {len(t)} | setup {med(1):5.2f} first-TMA {med(2):5.2f} "
                    f"first-stage-landed {med(3):5.2f} |")
            for i in range(min(tiles, 5)):
                line += f" tile{i}: mma-done {med(4 + 2 * i):5.2f} epi-done {med(5 + 2 * i):5.2f} |"
            line += f" exit {med(15):5.2f} (max {rel[:, 15].max():5.2f})"
            print(line)


if __name__ == "__main__":
    main()
