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
print("  tile 0 chunk 2: softmax hands it over %.2f | MMA thread sees it %.2f | PV MMAs issued %.2f | MMA thread sees chunk 3 %.2f" % tuple(med(i) for i in (27, 28, 29, 30)))

# ---- backward kernels (debug hook srw_attn_set_bwd_trace): per-CTA stamps of the MMA thread and of element-wise thread 0 ----
lib.srw_attn_set_bwd_trace.argtypes = [C.c_void_p]
Bb = max(2, B * 2 // 3)                      # the backward runs on the gradient rows only (2/3 of the forward batch)
qkv_b = O.split_planes(torch.randn(Bb * N, 3 * H * 64, device="cuda"))
o_b, lse_b = O.attn_fwd(qkv_b, Bb, N, H)
do_b = O.split_planes(torch.randn(Bb * N, H * 64, device="cuda"))
for _ in range(3):
    O.attn_bwd(qkv_b, o_b, do_b, lse_b, Bb, N, H)
RT = (N + 127) // 128
ncta = Bb * H * RT
trace = torch.zeros(2 * ncta * 48, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
lib.srw_attn_set_bwd_trace(trace.data_ptr())
O.attn_bwd(qkv_b, o_b, do_b, lse_b, Bb, N, H)
torch.cuda.synchronize()
lib.srw_attn_set_bwd_trace(None)
tb = trace.cpu().numpy().reshape(2, ncta, 48).astype(np.float64)
nch = min(5, (N + 63) // 64)
for k, name in enumerate(("dQ", "dK/dV")):
    t = tb[k]
    rel = (t - t[:, :1]) / 1.9e3
    # full row tiles only (the last tile of N = 257 holds one row) and, separately, CTAs by start order
    full = np.arange(ncta) % RT < max(1, RT - 1)
    med = lambda c: float(np.median(rel[full, c]))
    start = t[:, 0] - t[:, 0].min()
    print(f"attn_bwd {name} B={Bb} N={N} H={H}: {ncta} CTAs, first-wave CTA span (entry -> stored) median {med(43):.2f} us; "
          f"kernel span {float((t[:, 43].max() - t[:, 0].min()) / 1.9e3):.2f} us; CTA start times: {np.sum(start / 1.9e3 < 2.0)} within 2 us")
    print(f"  R tiles landed {med(1):.2f}")
    for j in range(nch):
        print(f"  chunk {j}: MMA thread: C landed {med(2 + 4 * j):6.2f} | T(j+1) issued {med(3 + 4 * j):6.2f} | sees X {med(4 + 4 * j):6.2f} | Acc issued {med(5 + 4 * j):6.2f}"
              f"   || EW: T ready {med(22 + 4 * j):6.2f} | X computed {med(23 + 4 * j):6.2f} | handed over {med(25 + 4 * j):6.2f}")
    print(f"  accumulators complete {med(42):.2f} | stored {med(43):.2f}")
