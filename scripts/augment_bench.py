"""Input-pipeline measurement: srw_augment_batch on the device vs the PIL / torchvision pipeline the
reference runs per sample on host cores (cifar.py:34-49; one process, like a DataLoader worker).

    python scripts/augment_bench.py [--batch 4096] [--reps 50]

Prints one JSON line: device images/s (CUDA events on the launching stream, decisions already on the device = kernel only, and
end to end incl. drawing + packing + uploading the decisions on the host), algorithmic HBM bytes (3 B read + 12 B written per
pixel) against MEASURED_PEAKS.json, and the host pipeline's images/s on a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


MEAN = [x / 255 for x in [129.3, 124.1, 112.4]]     # cifar.py:18-21
STD = [x / 255 for x in [68.2, 65.4, 70.4]]


def host_pipeline(size=32):
    """CPU baseline leg (the only part of this script that touches oracle/ and tests/): the reference's transform_strong rebuilt from
    the same library calls (torchvision transforms + PIL ops in RandAugment order)."""
    from PIL import Image
    from torchvision import transforms
    from oracle import augment_oracle as A
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_augment_oracle import _pil_op

    class RandAug:
        def __call__(self, img):
            ops, cut = A.draw_randaugment(size)
            for op, v in ops:
                img = _pil_op(op, img, v)
            if cut is not None:
                from PIL import ImageDraw
                img = img.copy()
                ImageDraw.Draw(img).rectangle(cut, A.CUTOUT_COLOR)
            return img
    geo = [transforms.Resize(size), transforms.RandomCrop(size, padding=int(size * 0.125), padding_mode="reflect"), transforms.RandomHorizontalFlip()]
    norm = [transforms.ToTensor(), transforms.Normalize(A.CIFAR100_MEAN, A.CIFAR100_STD)]
    return transforms.Compose(geo + norm), transforms.Compose(geo + [RandAug()] + norm), Image


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--cpu-images", type=int, default=2000)
    a = ap.parse_args()
    from semireward_b200 import _lib as L
    from semireward_b200.datasets import gpu_augment as G
    S = a.size
    rng = np.random.default_rng(0)
    data = rng.integers(0, 256, (50000, S, S, 3), dtype=np.uint8)
    pipe = G.DeviceImagePipeline(data, MEAN, STD)
    torch.manual_seed(0); random.seed(0); np.random.seed(0)
    n = a.batch
    idx = rng.integers(0, 50000, n).tolist()
    # ---- end to end: draw + pack + upload + kernel (weak and strong view of every sample, like the unlabelled loader) ----
    for _ in range(3):
        pipe.weak_and_strong(idx)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        pipe.weak_and_strong(idx)
    torch.cuda.synchronize()
    e2e = 5 * 2 * n / (time.perf_counter() - t0)
    # ---- the same with the decisions drawn in bulk (draw_records: same distributions, one numpy Generator) ----
    g = np.random.default_rng(0)
    for _ in range(3):
        pipe.weak_and_strong_fast(idx, g)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        pipe.weak_and_strong_fast(idx, g)
    torch.cuda.synchronize()
    e2e_fast = 5 * 2 * n / (time.perf_counter() - t0)
    # ---- kernel only: records resident ----
    decs = [G.draw_strong(S, pipe.padding) for _ in idx]
    recs = pipe._upload(G.pack_samples(idx, decs, S))
    out = torch.empty(n, 3, S, S, device="cuda")
    args = L.AugmentArgs(src=L.ptr(pipe.data), n_src=50000, img_size=S, padding=pipe.padding, samples=L.ptr(recs), n=n,
                         mean=(L.f32 * 3)(*pipe.mean), std=(L.f32 * 3)(*pipe.std), out=L.ptr(out), out_u8=None)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.reps)]
    for r in range(a.reps + 3):
        flush.zero_()                                   # L2 flush between timed launches
        if r >= 3:
            ev[r - 3][0].record()
        L.check(pipe.lib.srw_augment_batch(C.byref(args), L.stream_ptr()))
        if r >= 3:
            ev[r - 3][1].record()
    torch.cuda.synchronize()
    ms = float(np.median([e0.elapsed_time(e1) for e0, e1 in ev]))
    bytes_alg = n * S * S * 15.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = None
    for k in ("hbm_gbs", "hbm_gbps"):
        if isinstance(peaks.get(k), (int, float)):
            peak = float(peaks[k]); break
    # ---- host pipeline on one core ----
    weak, strong, Image = host_pipeline(S)
    m = a.cpu_images
    t0 = time.perf_counter()
    for i in range(m):
        img = Image.fromarray(data[i])
        weak(img); strong(img)
    cpu = 2 * m / (time.perf_counter() - t0)
    print(json.dumps({"workload": f"transform_weak + transform_strong, {S}x{S}x3 uint8, batch {n} samples x 2 views",
                      "device_kernel_images_per_s": n / (ms * 1e-3), "kernel_ms": ms,
                      "roofline": {"bound": "hbm", "achieved": bytes_alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": (bytes_alg / (ms * 1e-3) / 1e9 / peak) if peak else None, "bytes_per_pixel": 15},
                      "device_e2e_images_per_s": e2e, "device_e2e_bulk_draw_images_per_s": e2e_fast, "host_pipeline_images_per_s_1core": cpu, "host_sample_images": 2 * m}))


if __name__ == "__main__":
    main()
