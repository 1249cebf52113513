"""srw_augment_batch (csrc/srw_augment.cu, through the C ABI) against oracle/augment_oracle.py and the golden tensors of the
imported reference: uint8 images and normalised fp32 tensors must be BIT-identical (byte / integer work)."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import augment_oracle as A  # noqa: E402


def _pipe(images, **kw):
    from semireward_b200.datasets import DeviceImagePipeline
    return DeviceImagePipeline(images, A.CIFAR100_MEAN, A.CIFAR100_STD, **kw)


def _to_product(d):
    from semireward_b200.datasets import AugDecision
    return AugDecision(d.crop_top, d.crop_left, d.flip, list(d.ops), d.cutout)


def _images(rng, n, size):
    from test_augment_oracle import _images as mk
    return np.stack(mk(rng, n, size))


def test_golden_tensors_of_the_imported_reference():
    from golden_cases import augment_decisions_from_arrays
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "augment_cifar.npz"))
    pipe = _pipe(z["images"])
    n = len(z["images"])
    strong = [_to_product(d) for d in augment_decisions_from_arrays(z)]
    got = pipe.transform(range(n), strong).cpu().numpy()
    assert np.array_equal(got, z["strong"])
    weak = [A.Decision(int(g[0]), int(g[1]), bool(g[2])) for g in z["weak_geo"]]
    got = pipe.transform(range(n), [_to_product(d) for d in weak]).cpu().numpy()
    assert np.array_equal(got, z["weak"])


@pytest.mark.parametrize("size", [32, 96, 128])
def test_every_op_bit_exact_against_the_oracle(size):
    rng = np.random.default_rng(size)
    imgs = _images(rng, 24, size)
    pipe = _pipe(imgs)
    pad = pipe.padding
    idx, decs = [], []
    for op in range(14):
        lo, hi = A.OP_RANGE[op]
        for rep in range(24 if size == 32 else 8):
            v = lo + (hi - lo) * float(rng.random()) if rep else lo
            cut = None
            if rep % 3:
                s = float(rng.random()) * size * 0.5
                x0 = int(max(0, float(rng.random()) * size - s / 2)); y0 = int(max(0, float(rng.random()) * size - s / 2))
                cut = (x0, y0, min(size, x0 + s), min(size, y0 + s))
            decs.append(A.Decision(int(rng.integers(0, 2 * pad + 1)), int(rng.integers(0, 2 * pad + 1)), bool(rng.integers(0, 2)), [(op, v)], cut))
            idx.append(int(rng.integers(0, len(imgs))))
    # three-op chains as RandAugment(3, 5) draws them
    for rep in range(96 if size == 32 else 24):
        ops = []
        for op in rng.integers(0, 14, 3):
            lo, hi = A.OP_RANGE[int(op)]
            ops.append((int(op), lo + (hi - lo) * float(rng.random())))
        decs.append(A.Decision(int(rng.integers(0, 2 * pad + 1)), int(rng.integers(0, 2 * pad + 1)), bool(rng.integers(0, 2)), ops, None))
        idx.append(int(rng.integers(0, len(imgs))))
    out, u8 = pipe.transform(idx, [_to_product(d) for d in decs], return_u8=True)
    out, u8 = out.cpu().numpy(), u8.cpu().numpy()
    for k, (i, d) in enumerate(zip(idx, decs)):
        want8 = A.transform_u8(imgs[i], d, pad)
        assert np.array_equal(u8[k], want8), (k, d, int((u8[k] != want8).sum()))
        assert np.array_equal(out[k], A.to_tensor_normalize(want8, A.CIFAR100_MEAN, A.CIFAR100_STD)), (k, d)


def test_edge_cases():
    rng = np.random.default_rng(9)
    imgs = _images(rng, 6, 32)
    # no padding (crop_ratio 1): RandomCrop draws nothing and copies; extrapolating blend factors; rotate by a multiple of 360; single sample
    pipe = _pipe(imgs, crop_ratio=1.0)
    assert pipe.padding == 0
    decs = [A.Decision(0, 0, True, [(A.SHARPNESS, 1.7), (A.CONTRAST, 1.9), (A.COLOR, 1.3)], None),
            A.Decision(0, 0, False, [(A.ROTATE, 360.0), (A.BRIGHTNESS, 1.5), (A.ROTATE, -720.0)], (0, 0, 32, 32)),
            A.Decision(0, 0, False, [(A.SOLARIZE, 0.0), (A.POSTERIZE, 1.0), (A.SOLARIZE, 256.0)], (31, 31, 31.9, 31.2)),
            A.Decision(0, 0, False, [(A.TRANSLATE_X, 1.5), (A.TRANSLATE_Y, -1.5), (A.SHEAR_X, 40.0)], None)]
    out, u8 = pipe.transform([0, 1, 2, 3], [_to_product(d) for d in decs], return_u8=True)
    for k, d in enumerate(decs):
        assert np.array_equal(u8[k].cpu().numpy(), A.transform_u8(imgs[k], d, 0)), k
    one = pipe.transform([5], [_to_product(decs[0])]).cpu().numpy()
    assert np.array_equal(one[0], A.transform(imgs[5], decs[0], 0, A.CIFAR100_MEAN, A.CIFAR100_STD))
    with pytest.raises(IndexError):
        pipe.transform([6], [_to_product(decs[0])])
    with pytest.raises(ValueError):
        pipe.transform([], [])
    with pytest.raises(ValueError):
        pipe.transform([0], [_to_product(A.Decision(1, 0, False))])        # offset outside the (unpadded) image


def test_loader_draws_the_reference_decision_stream():
    """DeviceSSLLoader under seeded generators = the oracle's draws (pinned to the reference's transforms) in __getitem__ order."""
    from semireward_b200.datasets import DeviceSSLLoader
    rng = np.random.default_rng(4)
    lb_imgs, ulb_imgs = _images(rng, 12, 32), _images(rng, 40, 32)
    targets = rng.integers(0, 100, 12)
    lb_batches = [[3, 1, 7, 0], [2, 2, 11, 5]]
    ulb_batches = [[9, 30, 1, 17, 5, 6, 39, 0], [8, 8, 2, 3, 21, 22, 23, 24]]
    loader = DeviceSSLLoader(_pipe(lb_imgs), targets, _pipe(ulb_imgs), lb_batches, ulb_batches)
    torch.manual_seed(7); random.seed(7); np.random.seed(7)
    got = [(dict((k, v.cpu().numpy()) for k, v in a.items()), dict((k, v.cpu().numpy()) for k, v in b.items())) for a, b in loader]
    torch.manual_seed(7); random.seed(7); np.random.seed(7)
    for (g_lb, g_ulb), ib, iu in zip(got, lb_batches, ulb_batches):
        x_lb = np.stack([A.transform(lb_imgs[i], A.draw_weak(32, 4), 4, A.CIFAR100_MEAN, A.CIFAR100_STD) for i in ib])
        xw, xs = [], []
        for i in iu:
            xw.append(A.transform(ulb_imgs[i], A.draw_weak(32, 4), 4, A.CIFAR100_MEAN, A.CIFAR100_STD))
            xs.append(A.transform(ulb_imgs[i], A.draw_strong(32, 4), 4, A.CIFAR100_MEAN, A.CIFAR100_STD))
        assert np.array_equal(g_lb["x_lb"], x_lb) and np.array_equal(g_lb["y_lb"], targets[ib]) and np.array_equal(g_lb["idx_lb"], ib)
        assert np.array_equal(g_ulb["x_ulb_w"], np.stack(xw)) and np.array_equal(g_ulb["x_ulb_s"], np.stack(xs))
        assert np.array_equal(g_ulb["idx_ulb"], iu)


def test_val_transform_and_train_steps_from_the_raw_dataset():
    """End to end on the device: uint8 dataset -> srw_augment_batch (weak / strong views drawn like the reference's loaders) -> native
    SRFlexMatch steps, against the oracle stepping on the ORACLE's transforms of the same images under the same seeds."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import build_native, build_oracle, small_cfg
    from semireward_b200.datasets import DeviceSSLLoader
    rng = np.random.default_rng(21)
    lb_imgs, ulb_imgs = _images(rng, 16, 32), _images(rng, 64, 32)
    targets = rng.integers(0, 100, 16)
    pipe_lb, pipe_ulb = _pipe(lb_imgs), _pipe(ulb_imgs)
    val = pipe_lb.val(range(4)).cpu().numpy()
    for i in range(4):
        assert np.array_equal(val[i], A.to_tensor_normalize(lb_imgs[i], A.CIFAR100_MEAN, A.CIFAR100_STD))
    cfg = small_cfg()
    orc, alg = build_oracle(cfg, 2), build_native(cfg, 2)
    lb_batches = [list(rng.integers(0, 16, 8)) for _ in range(3)]
    ulb_batches = [list(rng.choice(64, 8, replace=False)) for _ in range(3)]
    torch.manual_seed(5); random.seed(5); np.random.seed(5)
    native_batches = list(DeviceSSLLoader(pipe_lb, targets, pipe_ulb, lb_batches, ulb_batches))
    torch.manual_seed(5); random.seed(5); np.random.seed(5)
    for it, ((d_lb, d_ulb), ib, iu) in enumerate(zip(native_batches, lb_batches, ulb_batches)):
        x_lb = np.stack([A.transform(lb_imgs[i], A.draw_weak(32, 4), 4, A.CIFAR100_MEAN, A.CIFAR100_STD) for i in ib])
        xw, xs = [], []
        for i in iu:
            xw.append(A.transform(ulb_imgs[i], A.draw_weak(32, 4), 4, A.CIFAR100_MEAN, A.CIFAR100_STD))
            xs.append(A.transform(ulb_imgs[i], A.draw_strong(32, 4), 4, A.CIFAR100_MEAN, A.CIFAR100_STD))
        batch = dict(x_lb=torch.from_numpy(x_lb), y_lb=torch.from_numpy(targets[ib]), idx_ulb=torch.tensor([int(i) for i in iu]),
                     x_ulb_w=torch.from_numpy(np.stack(xw)), x_ulb_s=torch.from_numpy(np.stack(xs)))
        rec = orc.train_step(dict(batch), it)
        orc.param_update()
        alg.it = it
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**d_lb, **d_ulb))
        alg.call_hook("after_train_step")
        torch.cuda.synchronize()
        assert abs(alg.log_dict["train/total_loss"] - float(rec["total_loss"])) < 1e-3
        assert torch.equal(alg._last_mask.cpu(), rec["mask"]) and torch.equal(alg._last_pseudo_label.cpu(), rec["pseudo"])


def test_record_array_route_equals_the_per_sample_route():
    """transform_records (numpy records, the throughput route) launches the same kernel on the same bytes as transform()."""
    from semireward_b200.datasets import draw_records, records_from_decisions
    rng = np.random.default_rng(33)
    imgs = _images(rng, 24, 32)
    pipe = _pipe(imgs)
    idx = [int(i) for i in rng.integers(0, 24, 96)]
    torch.manual_seed(2); random.seed(2); np.random.seed(2)
    decs = [A.draw_strong(32, 4) for _ in idx]
    prod = [_to_product(d) for d in decs]
    a = pipe.transform(idx, prod).cpu().numpy()
    b = pipe.transform_records(records_from_decisions(idx, prod, 32)).cpu().numpy()
    assert np.array_equal(a, b)
    for k, (i, d) in enumerate(zip(idx, decs)):
        assert np.array_equal(b[k], A.transform(imgs[i], d, 4, A.CIFAR100_MEAN, A.CIFAR100_STD))
    # bulk-drawn records: a finite, normalised batch of the right shape for both views
    w, s = pipe.weak_and_strong_fast(idx, np.random.default_rng(5))
    assert w.shape == s.shape == (96, 3, 32, 32) and torch.isfinite(w).all() and torch.isfinite(s).all()
    rec = draw_records(idx, 32, 4, False, np.random.default_rng(6))
    got = pipe.transform_records(rec).cpu().numpy()
    for k in range(8):
        d = A.Decision(int(rec["crop_top"][k]), int(rec["crop_left"][k]), bool(rec["flip"][k]))
        assert np.array_equal(got[k], A.transform(imgs[idx[k]], d, 4, A.CIFAR100_MEAN, A.CIFAR100_STD))
