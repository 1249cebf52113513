"""Per-CTA clock64 timeline of attn_fwd_kernel (debug hook srw_attn_set_trace): python scripts/attn_trace.py [--B 24]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semireward_b200 import _lib as L, ops as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=24)
ap.add_argument("--N", type=int, default=257)
ap.add_argument("--H", type=int, default=6)
a = ap.parse_args()
lib = L.load()
lib.srw_attn_set_trace.argtypes = [C.c_void_p]
B, N, H = a.B, a.N, a.H
qkv = O.split_planes(torch.randn(B * N, 3 * H * 64, device="cuda"))
for _ in range(3):
    O.attn_fwd(qkv, B, N, H)
trace = torch.zeros(B * H * 32, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
lib.srw_attn_set_trace(trace.data_ptr())
O.attn_fwd(qkv, B, N, H)
torch.cuda.synchronize()
lib.srw_attn_set_trace(None)
t = trace.cpu().numpy().reshape(B * H, 32).astype(np.float64)
rel = (t - t[:, :1]) / 1.9e3
med = lambda c: float(np.median(rel[:, c]))
print(f"attn_fwd B={B} N={N} H={H} (us since CTA entry, median over {B * H} CTAs): K/V landed {med(1):.2f}")
names = ["S ready", "row max", "P chunk 0", "P last", "O done", "stored"]
for tl in range((N + 127) // 128):
    print(f"  tile {tl}: " + " | ".join(f"{n} {med(2 + 8 * tl + i):6.2f}" for i, n in enumerate(names)))
print("  tile 0 chunk loop: softmax sees buffer free (c=2) %.2f | hands chunk 2 over %.2f | MMA thread sees chunk 2 %.2f | chunk 2 issued+committed %.2f | "
      "MMA thread sees chunk 3 %.2f | softmax sees chunk 2's buffer free (c=4) %.2f" % tuple(med(i) for i in (26, 27, 28, 29, 30, 31)))
