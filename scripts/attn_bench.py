"""Microbenchmark of srw_attn_fwd / srw_attn_bwd at the ViT-S step's shape (B images x 6 heads x 257 tokens)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semireward_b200 import ops as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--B", type=int, default=24)
ap.add_argument("--N", type=int, default=257)
ap.add_argument("--H", type=int, default=6)
a = ap.parse_args()
B, N, H = a.B, a.N, a.H
D = H * 64
torch.manual_seed(0)
qkv = O.split_planes(torch.randn(B * N, 3 * D, device="cuda"))
d_o = O.split_planes(torch.randn(B * N, D, device="cuda"))
o, lse = O.attn_fwd(qkv, B, N, H)
for name, fn in (("attn_fwd", lambda: O.attn_fwd(qkv, B, N, H)), ("attn_bwd", lambda: O.attn_bwd(qkv, o, d_o, lse, B, N, H))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / a.reps
    fl = (4.0 if name == "attn_fwd" else 8.0) * B * H * N * N * 64
    print(f"{name} B={B} N={N} H={H}: {us:8.1f} us  {fl / (us * 1e-6) / 1e12:6.1f} TFLOP/s algorithmic")
