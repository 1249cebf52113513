"""Pins oracle/augment_oracle.py (numpy restatement of the reference's image pipeline) to the live libraries the reference
calls — Pillow and torchvision, both installed on the build container AND on the GPU box — and, where /root/reference is
mounted, to the reference's own RandAugment / transform objects under seeded generators.  CPU only."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import augment_oracle as A

PIL = pytest.importorskip("PIL")
from PIL import Image, ImageDraw, ImageEnhance, ImageFilter, ImageOps  # noqa: E402

HAVE_REF = os.path.isdir("/root/reference/semilearn")


def _images(rng, n, size=32):
    """random images plus the degenerate ones the histogram ops branch on."""
    out = []
    for k in range(n):
        kind = k % 6
        if kind == 0:
            a = rng.integers(0, 256, (size, size, 3))
        elif kind == 1:                                   # narrow range (autocontrast stretches, equalize has few levels)
            lo = int(rng.integers(0, 200))
            a = rng.integers(lo, lo + int(rng.integers(2, 56)), (size, size, 3))
        elif kind == 2:                                   # smooth gradient + noise
            y, x = np.mgrid[0:size, 0:size]
            a = (y[..., None] * rng.integers(1, 8, 3) + x[..., None] * rng.integers(1, 8, 3) + rng.integers(0, 16, (size, size, 3))) % 256
        elif kind == 3:                                   # one constant channel, one two-level channel
            a = rng.integers(0, 256, (size, size, 3))
            a[..., 0] = int(rng.integers(0, 256))
            a[..., 1] = np.where(rng.random((size, size)) < 0.5, 3, 250)
        elif kind == 4:                                   # constant image
            a = np.zeros((size, size, 3), dtype=np.int64) + rng.integers(0, 256, 3)
        else:                                             # a few levels, very unequal counts (equalize step == 0 / small)
            a = rng.choice([0, 1, 128, 255], size=(size, size, 3), p=[0.9, 0.05, 0.03, 0.02])
        out.append(a.astype(np.uint8))
    return out


def _pil_op(op, img, v):
    w, h = img.size
    if op == A.AUTOCONTRAST:
        return ImageOps.autocontrast(img)
    if op == A.BRIGHTNESS:
        return ImageEnhance.Brightness(img).enhance(v)
    if op == A.COLOR:
        return ImageEnhance.Color(img).enhance(v)
    if op == A.CONTRAST:
        return ImageEnhance.Contrast(img).enhance(v)
    if op == A.EQUALIZE:
        return ImageOps.equalize(img)
    if op == A.IDENTITY:
        return img
    if op == A.POSTERIZE:
        return ImageOps.posterize(img, max(1, int(v)))
    if op == A.ROTATE:
        return img.rotate(v)
    if op == A.SHARPNESS:
        return ImageEnhance.Sharpness(img).enhance(v)
    if op == A.SHEAR_X:
        return img.transform(img.size, Image.AFFINE, (1, v, 0, 0, 1, 0))
    if op == A.SHEAR_Y:
        return img.transform(img.size, Image.AFFINE, (1, 0, 0, v, 1, 0))
    if op == A.SOLARIZE:
        return ImageOps.solarize(img, v)
    if op == A.TRANSLATE_X:
        return img.transform(img.size, Image.AFFINE, (1, 0, v * w, 0, 1, 0))
    if op == A.TRANSLATE_Y:
        return img.transform(img.size, Image.AFFINE, (1, 0, 0, 0, 1, v * h))
    raise ValueError(op)


@pytest.mark.parametrize("op", range(14))
def test_every_op_matches_pillow_bit_for_bit(op):
    rng = np.random.default_rng(100 + op)
    lo, hi = A.OP_RANGE[op]
    n = 0
    for size in (32, 96):
        for a in _images(rng, 36 if size == 32 else 6, size):
            for v in [lo, lo + (hi - lo) * 0.5] + [lo + (hi - lo) * float(rng.random()) for _ in range(4)]:
                want = np.asarray(_pil_op(op, Image.fromarray(a), v))
                got = A.apply_op(a, op, v)
                assert np.array_equal(got, want), (A.OP_NAMES[op], v, int((got != want).sum()))
                n += 1
    assert n >= 200


def test_blend_extrapolating_branch_and_smooth_filter():
    rng = np.random.default_rng(7)
    for a in _images(rng, 12):
        img = Image.fromarray(a)
        assert np.array_equal(A.smooth(a), np.asarray(img.filter(ImageFilter.SMOOTH)))
        for v in (1.3, 1.9, 0.0, 1.0):
            assert np.array_equal(A.sharpness(a, v), np.asarray(ImageEnhance.Sharpness(img).enhance(v)))
            assert np.array_equal(A.contrast(a, v), np.asarray(ImageEnhance.Contrast(img).enhance(v)))


def test_rectangle_is_inclusive_and_truncates():
    rng = np.random.default_rng(3)
    for a in _images(rng, 6):
        for _ in range(20):
            v = float(rng.random()) * 16
            x0 = int(max(0, float(rng.random()) * 32 - v / 2))
            y0 = int(max(0, float(rng.random()) * 32 - v / 2))
            xy = (x0, y0, min(32, x0 + v), min(32, y0 + v))
            img = Image.fromarray(a).copy()
            ImageDraw.Draw(img).rectangle(xy, A.CUTOUT_COLOR)
            assert np.array_equal(A.cutout_abs(a, xy), np.asarray(img)), xy


def test_crop_flip_and_normalise_match_torchvision():
    from torchvision import transforms
    rng = np.random.default_rng(5)
    norm = transforms.Compose([transforms.ToTensor(), transforms.Normalize(A.CIFAR100_MEAN, A.CIFAR100_STD)])
    for k, a in enumerate(_images(rng, 12)):
        weak = transforms.Compose([transforms.Resize(32), transforms.RandomCrop(32, padding=4, padding_mode="reflect"),
                                   transforms.RandomHorizontalFlip()])
        torch.manual_seed(k)
        want = weak(Image.fromarray(a))
        torch.manual_seed(k)
        d = A.draw_weak(32, 4)
        got = A.transform_u8(a, d, 4)
        assert np.array_equal(got, np.asarray(want))
        assert np.array_equal(A.to_tensor_normalize(got, A.CIFAR100_MEAN, A.CIFAR100_STD), norm(want).numpy())   # bit-exact floats


@pytest.mark.skipif(not HAVE_REF, reason="live reference only exists in the build container")
def test_strong_transform_matches_the_live_reference_under_seeded_generators():
    """The reference's own transform_strong (cifar.py:42-49) with its RandAugment, seeded; the oracle draws the same decisions
    from the same generators and must reproduce the tensor bit for bit."""
    from oracle import ref_driver as R
    R.load_reference()
    from torchvision import transforms
    from semilearn.datasets.augmentation import RandAugment
    strong = transforms.Compose([transforms.Resize(32), transforms.RandomCrop(32, padding=4, padding_mode="reflect"),
                                 transforms.RandomHorizontalFlip(), RandAugment(3, 5), transforms.ToTensor(),
                                 transforms.Normalize(A.CIFAR100_MEAN, A.CIFAR100_STD)])
    rng = np.random.default_rng(11)
    seen = set()
    for k, a in enumerate(_images(rng, 240)):
        torch.manual_seed(k); random.seed(k); np.random.seed(k)
        want = strong(Image.fromarray(a)).numpy()
        torch.manual_seed(k); random.seed(k); np.random.seed(k)
        d = A.draw_strong(32, 4)
        seen.update(op for op, _ in d.ops)
        got = A.transform(a, d, 4, A.CIFAR100_MEAN, A.CIFAR100_STD)
        assert np.array_equal(got, want), (k, d)
    assert len(seen) == 14


def test_golden_fixture():
    """tests/golden/augment_cifar.npz: outputs of the imported reference (make_golden_augment.py), decisions included."""
    path = os.path.join(os.path.dirname(__file__), "golden", "augment_cifar.npz")
    z = np.load(path)
    from golden_cases import augment_decisions_from_arrays
    imgs, want = z["images"], z["strong"]
    decs = augment_decisions_from_arrays(z)
    for i in range(len(imgs)):
        got = A.transform(imgs[i], decs[i], 4, A.CIFAR100_MEAN, A.CIFAR100_STD)
        assert np.array_equal(got, want[i]), i


def test_host_records_carry_what_pillows_python_layer_computes():
    """semireward_b200.datasets.gpu_augment (product host code, no GPU needed): the decision stream equals the oracle's under the
    same seeds, and pack_samples turns a decision into the record fields the kernel consumes — C-float blend factor, posterize
    bits, ceil of the solarize threshold, Pillow's affine coefficients, ImageDraw's truncated rectangle."""
    import ctypes as C
    import math
    from semireward_b200 import _lib as L
    from semireward_b200.datasets import gpu_augment as G
    assert C.sizeof(L.AugSample) == 232 and C.sizeof(L.AugOpDesc) == 64
    for k in range(100):
        torch.manual_seed(k); random.seed(k); np.random.seed(k)
        a = [G.draw_weak(32, 4), G.draw_strong(32, 4)]
        torch.manual_seed(k); random.seed(k); np.random.seed(k)
        b = [A.draw_weak(32, 4), A.draw_strong(32, 4)]
        for x, y in zip(a, b):
            assert (x.crop_top, x.crop_left, x.flip, x.ops, x.cutout) == (y.crop_top, y.crop_left, y.flip, y.ops, y.cutout)
        rec = G.pack_samples([7, 9], a, 32)
        assert rec[0].n_ops == 0 and rec[0].src_index == 7 and rec[0].cut_x1 < rec[0].cut_x0
        s = rec[1]
        assert (s.src_index, s.crop_top, s.crop_left, s.flip, s.n_ops) == (9, a[1].crop_top, a[1].crop_left, int(a[1].flip), 3)
        assert (s.cut_x0, s.cut_y0, s.cut_x1, s.cut_y1) == tuple(int(v) for v in b[1].cutout)
        for j, (op, v) in enumerate(b[1].ops):
            o = s.ops[j]
            assert o.op == op
            if op in (A.BRIGHTNESS, A.COLOR, A.CONTRAST, A.SHARPNESS):
                assert o.alpha == float(np.float32(v))
            elif op == A.POSTERIZE:
                assert o.ival == max(1, int(v))
            elif op == A.SOLARIZE:
                assert o.ival == math.ceil(v) and all((i < v) == (i < o.ival) for i in range(257))
            elif op in (A.ROTATE, A.SHEAR_X, A.SHEAR_Y, A.TRANSLATE_X, A.TRANSLATE_Y):
                m = A.op_matrix(op, v, 32, 32)
                assert (o.identity == 1) if m is None else list(o.a) == [float(t) for t in m]
    # no-colour list of RandAugment(exclude_color_aug=True) (randaugment.py:176-193)
    random.seed(3)
    assert all(op in (1, 4, 5, 7, 8, 9, 10, 12, 13) for _ in range(50) for op, _v in G.draw_strong(32, 4, exclude_color_aug=True).ops)
    with pytest.raises(ValueError):
        G.pack_samples([0], [G.AugDecision(0, 0, False, [(A.SOLARIZE, 300.0)])], 32)
    with pytest.raises(ValueError):
        G.pack_samples([0], [G.AugDecision(0, 0, False, [(A.IDENTITY, 0.0)] * 4)], 32)


def test_vectorised_records_equal_the_per_sample_packing_byte_for_byte():
    """pack_arrays / draw_records (the throughput route of the input pipeline) against pack_samples, the route the GPU parity tests
    validate: same bytes for the same decisions, every op, degenerate magnitudes included; layout of the numpy mirror == ctypes."""
    import ctypes as C
    from semireward_b200 import _lib as L
    from semireward_b200.datasets import gpu_augment as G
    for name in G.SAMPLE_DTYPE.names:
        assert G.SAMPLE_DTYPE.fields[name][1] == getattr(L.AugSample, {"cut": "cut_x0"}.get(name, name)).offset, name
    for name in G.OP_DTYPE.names:
        assert G.OP_DTYPE.fields[name][1] == getattr(L.AugOpDesc, name).offset, name
    assert G.SAMPLE_DTYPE.itemsize == C.sizeof(L.AugSample) == 232
    rng = np.random.default_rng(12)
    for size, pad in ((32, 4), (96, 12)):
        n = 600
        idx = rng.integers(0, 1000, n)
        top, left, flip = rng.integers(0, 2 * pad + 1, n), rng.integers(0, 2 * pad + 1, n), rng.random(n) < 0.5
        ops = rng.integers(0, 14, (n, 3))
        lo = np.array([r[0] for r in A.OP_RANGE], dtype=np.float64)[ops]
        hi = np.array([r[1] for r in A.OP_RANGE], dtype=np.float64)[ops]
        val = lo + (hi - lo) * rng.random((n, 3))
        val[:14, 0], ops[:14, 0] = lo[:14, 0] * 0 + np.array([r[0] for r in A.OP_RANGE]), np.arange(14)      # every op at its lower end
        val[14, :], ops[14, :] = [360.0, -720.0, 0.0], [A.ROTATE] * 3                                        # rotations Image.rotate copies
        v = rng.random(n) * 0.5 * size
        x0 = np.maximum(0.0, rng.random(n) * size - v / 2).astype(np.int64).astype(np.float64)
        y0 = np.maximum(0.0, rng.random(n) * size - v / 2).astype(np.int64).astype(np.float64)
        cut = np.stack([x0, y0, np.minimum(float(size), x0 + v), np.minimum(float(size), y0 + v)], axis=1)
        cut[::7] = np.nan
        decs = [G.AugDecision(int(top[i]), int(left[i]), bool(flip[i]), [(int(o), float(x)) for o, x in zip(ops[i], val[i])],
                              None if np.isnan(cut[i, 0]) else tuple(float(c) for c in cut[i])) for i in range(n)]
        want = G.records_from_decisions(idx, decs, size)
        got = G.pack_arrays(idx, top, left, flip, ops, val, cut, size)
        assert got.tobytes() == want.tobytes()
        weak = [G.AugDecision(int(top[i]), int(left[i]), bool(flip[i])) for i in range(n)]
        assert G.pack_arrays(idx, top, left, flip, size=size).tobytes() == G.records_from_decisions(idx, weak, size).tobytes()
    # the bulk drawer: distributions of RandAugment(3, 5) / RandomCrop / flip / Cutout
    rec = G.draw_records(np.arange(20000) % 500, 32, 4, True, np.random.default_rng(1))
    assert rec["n_ops"].min() == 3 and set(np.unique(rec["ops"]["op"])) == set(range(14))
    assert abs(rec["flip"].mean() - 0.5) < 0.02 and rec["crop_top"].min() == 0 and rec["crop_top"].max() == 8 and abs(rec["crop_left"].mean() - 4) < 0.1
    counts = np.bincount(rec["ops"]["op"].ravel(), minlength=14) / rec["ops"]["op"].size
    assert np.abs(counts - 1 / 14).max() < 0.01
    w = rec["cut"][:, 2] - rec["cut"][:, 0]
    assert w.min() >= 0 and w.max() <= 16 and rec["cut"][:, 2].max() <= 32
    post = rec["ops"]["ival"][rec["ops"]["op"] == A.POSTERIZE]
    assert set(np.unique(post)) == {4, 5, 6, 7}
    weak = G.draw_records(np.arange(100), 32, 4, False, np.random.default_rng(2))
    assert weak["n_ops"].max() == 0 and (weak["cut"][:, 2] < weak["cut"][:, 0]).all()
    with pytest.raises(ValueError):
        G.pack_arrays([0], [0], [0], [False], np.array([[A.SOLARIZE]]), np.array([[300.0]]), None, 32)


def test_loader_glue_on_a_stub_pipeline():
    """DeviceSSLLoader's host glue without a device: batch keys, shapes, targets gather and which pipeline entry points each route
    calls (per-sample reference streams vs bulk records)."""
    from semireward_b200.datasets import gpu_augment as G

    class Stub:
        size, padding = 32, 4

        def __init__(self):
            self.data = torch.zeros(50, 32, 32, 3, dtype=torch.uint8)
            self.calls = []

        def weak(self, idx):
            self.calls.append(("weak", len(idx)))
            return torch.zeros(len(idx), 3, 32, 32)

        def weak_and_strong(self, idx):
            self.calls.append(("weak_and_strong", len(idx)))
            return torch.zeros(len(idx), 3, 32, 32), torch.ones(len(idx), 3, 32, 32)

        def transform_records(self, rec):
            assert rec.dtype == G.SAMPLE_DTYPE and rec["n_ops"].max() == 0
            self.calls.append(("records", len(rec)))
            return torch.zeros(len(rec), 3, 32, 32)

        def weak_and_strong_fast(self, idx, rng):
            self.calls.append(("fast", len(idx)))
            return torch.zeros(len(idx), 3, 32, 32), torch.ones(len(idx), 3, 32, 32)
    targets = np.arange(50) % 7
    for rng, want in ((None, [("weak", 4)], ), (np.random.default_rng(0), [("records", 4)])):
        lb, ulb = Stub(), Stub()
        out = list(G.DeviceSSLLoader(lb, targets, ulb, [[3, 9, 1, 1], [4, 5, 6, 7]], [[0, 1, 2, 3, 4, 5, 6, 7], [8, 9, 10, 11, 12, 13, 14, 15]], bulk_rng=rng))
        assert len(out) == 2 and lb.calls == want * 2 and ulb.calls == [("fast" if rng is not None else "weak_and_strong", 8)] * 2
        d_lb, d_ulb = out[0]
        assert set(d_lb) == {"idx_lb", "x_lb", "y_lb"} and set(d_ulb) == {"idx_ulb", "x_ulb_w", "x_ulb_s"}
        assert d_lb["y_lb"].tolist() == [3, 2, 1, 1] and d_lb["idx_lb"].tolist() == [3, 9, 1, 1] and d_ulb["idx_ulb"].tolist() == list(range(8))
        assert d_ulb["x_ulb_s"].shape == (8, 3, 32, 32) and d_lb["x_lb"].shape == (4, 3, 32, 32)
