"""Build container only: run the LIVE reference's strong / weak image transforms (semilearn/datasets/cv_datasets/cifar.py:34-49
with semilearn.datasets.augmentation.RandAugment) on seeded synthetic 32 x 32 images and store inputs, decisions and outputs in
tests/golden/augment_cifar.npz.  The decisions are recovered by replaying the same seeds through oracle.augment_oracle.draw_*
(the test asserts that replay reproduces the reference's tensors, so the stored decisions are the reference's).

    python tests/golden/make_golden_augment.py
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import augment_oracle as A   # noqa: E402
from oracle import ref_driver as R       # noqa: E402
from golden_cases import augment_decisions_to_arrays   # noqa: E402


def main():
    R.load_reference()
    from PIL import Image
    from torchvision import transforms
    from semilearn.datasets.augmentation import RandAugment
    norm = [transforms.ToTensor(), transforms.Normalize(A.CIFAR100_MEAN, A.CIFAR100_STD)]
    geo = [transforms.Resize(32), transforms.RandomCrop(32, padding=4, padding_mode="reflect"), transforms.RandomHorizontalFlip()]
    strong = transforms.Compose(geo + [RandAugment(3, 5)] + norm)
    weak = transforms.Compose(geo + norm)
    rng = np.random.default_rng(2024)
    n = 48
    y, x = np.mgrid[0:32, 0:32]
    imgs = []
    for k in range(n):
        if k % 3 == 0:
            a = rng.integers(0, 256, (32, 32, 3))
        elif k % 3 == 1:     # smooth structure: geometric ops and the histogram ops see something image-like
            a = (y[..., None] * rng.integers(1, 8, 3) + x[..., None] * rng.integers(1, 8, 3) + rng.integers(0, 24, (32, 32, 3))) % 256
        else:                # narrow range
            lo = int(rng.integers(0, 180))
            a = rng.integers(lo, lo + int(rng.integers(3, 70)), (32, 32, 3))
        imgs.append(a.astype(np.uint8))
    imgs = np.stack(imgs)
    out_s, out_w, decs = [], [], []
    for k in range(n):
        seed = 1000 + k
        torch.manual_seed(seed); random.seed(seed); np.random.seed(seed)
        pil = Image.fromarray(imgs[k])
        out_w.append(weak(pil).numpy())            # BasicDataset.__getitem__ order: weak first, then strong (datasetbase.py:90,115)
        out_s.append(strong(pil).numpy())
        torch.manual_seed(seed); random.seed(seed); np.random.seed(seed)
        dw = A.draw_weak(32, 4)
        ds = A.draw_strong(32, 4)
        decs.append((dw, ds))
    arr_s = augment_decisions_to_arrays([d[1] for d in decs])
    arr_w = augment_decisions_to_arrays([d[0] for d in decs])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "augment_cifar.npz")
    np.savez_compressed(path, images=imgs, strong=np.stack(out_s), weak=np.stack(out_w), weak_geo=arr_w["geo"], **arr_s)
    print(path, os.path.getsize(path), "bytes; ops seen:", sorted(set(int(o) for o in arr_s["op_id"].ravel())))


if __name__ == "__main__":
    main()
