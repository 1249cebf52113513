"""`transform_weak` / `transform_strong` of the reference's image datasets on the device.

Reference: get_cifar builds the two torchvision pipelines (semilearn/datasets/cv_datasets/cifar.py:34-49) and
BasicDataset.__getitem__ applies them per sample in DataLoader workers (semilearn/datasets/cv_datasets/datasetbase.py:74-115):
weak = Resize -> RandomCrop(reflect padding) -> RandomHorizontalFlip -> ToTensor -> Normalize; strong adds RandAugment(3, 5) + Cutout
(semilearn/datasets/augmentation/randaugment.py:157-206).  Here the uint8 array stays resident in HBM, the host only DRAWS the
random decisions — from the same generators, in the same order as the reference's objects (torch's global generator for crop /
flip, Python's `random` for ops and magnitudes, numpy's global generator for the Cutout position) — and `srw_augment_batch`
(csrc/srw_augment.cu) does all pixel work: one CTA per sample, bit-identical to Pillow / torchvision.

There is no CPU fallback: without the CUDA library the pipeline raises.
"""
from __future__ import annotations

import ctypes as C
import math
import random
from dataclasses import dataclass, field
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib as L

# positions in the reference's augment_list() (randaugment.py:157-174) = enum srw_aug_op
(AUTOCONTRAST, BRIGHTNESS, COLOR, CONTRAST, EQUALIZE, IDENTITY, POSTERIZE, ROTATE, SHARPNESS, SHEAR_X, SHEAR_Y, SOLARIZE,
 TRANSLATE_X, TRANSLATE_Y) = range(14)
_RANGE = [(0, 1), (0.05, 0.95), (0.05, 0.95), (0.05, 0.95), (0, 1), (0, 1), (4, 8), (-30, 30), (0.05, 0.95), (-0.3, 0.3), (-0.3, 0.3),
          (0, 256), (-0.3, 0.3), (-0.3, 0.3)]
_NO_COLOR = [BRIGHTNESS, EQUALIZE, IDENTITY, ROTATE, SHARPNESS, SHEAR_X, SHEAR_Y, TRANSLATE_X, TRANSLATE_Y]   # augment_list_no_color()


@dataclass
class AugDecision:
    """Everything random about one transformed sample."""
    crop_top: int = 0
    crop_left: int = 0
    flip: bool = False
    ops: List[Tuple[int, float]] = field(default_factory=list)       # (op id, val), application order; empty = weak transform
    cutout: Optional[Tuple[float, float, float, float]] = None       # the xy CutoutAbs hands to ImageDraw.rectangle


def _draw_geometry(size: int, padding: int) -> Tuple[int, int, bool]:
    # RandomCrop.get_params: no draw when the padded image equals the crop; RandomHorizontalFlip: torch.rand(1) < p
    top = left = 0
    if padding > 0:
        top = int(torch.randint(0, 2 * padding + 1, size=(1,)).item())
        left = int(torch.randint(0, 2 * padding + 1, size=(1,)).item())
    return top, left, bool(torch.rand(1) < 0.5)


def draw_weak(size: int, padding: int) -> AugDecision:
    return AugDecision(*_draw_geometry(size, padding))


def draw_strong(size: int, padding: int, n: int = 3, exclude_color_aug: bool = False) -> AugDecision:
    """RandAugment.__call__ (randaugment.py:196-203) after the geometric front: random.choices, one random.random() per op, one for
    the Cutout size, CutoutAbs' two np.random.uniform draws (randaugment.py:136-146)."""
    top, left, flip = _draw_geometry(size, padding)
    table = _NO_COLOR if exclude_color_aug else list(range(14))
    ops = []
    for i in random.choices(table, k=n):
        lo, hi = _RANGE[i]
        ops.append((i, lo + float(hi - lo) * random.random()))
    v = random.random() * 0.5
    cut = None
    if v > 0.0:
        v = v * size
        x0 = np.random.uniform(size)
        y0 = np.random.uniform(size)
        x0 = int(max(0, x0 - v / 2.0))
        y0 = int(max(0, y0 - v / 2.0))
        cut = (x0, y0, min(size, x0 + v), min(size, y0 + v))
    return AugDecision(top, left, flip, ops, cut)


def _affine_coefficients(op: int, v: float, size: int) -> Optional[List[float]]:
    """Inverse-map coefficients as Pillow's Python layer forms them (Image.rotate rounds cos / sin to 15 decimals and moves the
    centre (w/2, h/2); the shears / translations pass the tuples of randaugment.py:66-110 through unchanged)."""
    if op == ROTATE:
        angle = v % 360.0
        if angle == 0:
            return None
        if angle in (90, 180, 270):
            raise NotImplementedError("Image.rotate's transpose fast paths (val is a continuous draw)")
        c = size / 2.0
        a = -math.radians(angle)
        m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
        m[2] = m[0] * (-c) + m[1] * (-c) + m[2]
        m[5] = m[3] * (-c) + m[4] * (-c) + m[5]
        m[2] += c
        m[5] += c
        return m
    if op == SHEAR_X:
        return [1.0, v, 0.0, 0.0, 1.0, 0.0]
    if op == SHEAR_Y:
        return [1.0, 0.0, 0.0, v, 1.0, 0.0]
    if op == TRANSLATE_X:
        return [1.0, 0.0, v * size, 0.0, 1.0, 0.0]
    return [1.0, 0.0, 0.0, 0.0, 1.0, v * size]


def pack_samples(indices: Sequence[int], decisions: Sequence[AugDecision], size: int):
    """-> ctypes array of srw_aug_sample (include/srw.h)."""
    n = len(decisions)
    assert len(indices) == n
    arr = (L.AugSample * n)()
    for s, idx, d in zip(arr, indices, decisions):
        s.src_index = int(idx)
        s.crop_top, s.crop_left, s.flip = d.crop_top, d.crop_left, int(d.flip)
        if len(d.ops) > 3:
            raise ValueError("srw_aug_sample holds at most 3 ops (RandAugment(3, 5))")
        s.n_ops = len(d.ops)
        for k, (op, v) in enumerate(d.ops):
            o = s.ops[k]
            o.op = op
            if op in (BRIGHTNESS, COLOR, CONTRAST, SHARPNESS):
                if v < 0.0:
                    raise ValueError("enhancement factor must be >= 0 (randaugment.py:21,26,31,59)")
                o.alpha = v
            elif op == POSTERIZE:
                o.ival = max(1, int(v))                               # randaugment.py:46-49
                if o.ival > 8:
                    raise ValueError("posterize bits > 8")
            elif op == SOLARIZE:
                if not 0 <= v <= 256:
                    raise ValueError("solarize threshold outside [0, 256] (randaugment.py:114)")
                o.ival = int(math.ceil(v))                            # integer level i is kept iff i < v  <=>  i < ceil(v)
            elif op in (ROTATE, SHEAR_X, SHEAR_Y, TRANSLATE_X, TRANSLATE_Y):
                m = _affine_coefficients(op, v, size)
                if m is None:
                    o.identity = 1
                else:
                    # libImaging takes the 16.16 fixed-point path only while every corner maps inside +-32768
                    for x, y in ((0, 0), (size, size), (0, size), (size, 0)):
                        if not (abs(x * m[0] + y * m[1] + m[2]) < 32768.0 and abs(x * m[3] + y * m[4] + m[5]) < 32768.0):
                            raise ValueError("affine coefficients outside the fixed-point range of Image.transform")
                    for j in range(6):
                        o.a[j] = m[j]
        if d.cutout is None:
            s.cut_x0, s.cut_y0, s.cut_x1, s.cut_y1 = 0, 0, -1, -1
        else:
            s.cut_x0, s.cut_y0, s.cut_x1, s.cut_y1 = (int(v) for v in d.cutout)   # ImageDraw truncates the float corners
    return arr


# numpy mirror of srw_aug_op_desc / srw_aug_sample (include/srw.h): every field naturally aligned, so the packed dtype IS the C layout
OP_DTYPE = np.dtype([("op", "<i4"), ("ival", "<i4"), ("alpha", "<f4"), ("identity", "<i4"), ("a", "<f8", (6,))])
SAMPLE_DTYPE = np.dtype([("src_index", "<i8"), ("crop_top", "<i4"), ("crop_left", "<i4"), ("flip", "<i4"), ("n_ops", "<i4"),
                         ("ops", OP_DTYPE, (3,)), ("cut", "<i4", (4,))])
assert OP_DTYPE.itemsize == C.sizeof(L.AugOpDesc) and SAMPLE_DTYPE.itemsize == C.sizeof(L.AugSample)


def records_from_decisions(indices: Sequence[int], decisions: Sequence[AugDecision], size: int) -> np.ndarray:
    """pack_samples as a numpy structured array (same bytes)."""
    return np.frombuffer(bytes(memoryview(pack_samples(indices, decisions, size)).cast("B")), dtype=SAMPLE_DTYPE).copy()


def pack_arrays(idx, top, left, flip, ops=None, val=None, cut=None, size: int = 32) -> np.ndarray:
    """pack_samples over arrays: idx / top / left / flip [n]; ops int [n, k <= 3] and val float64 [n, k] (None = weak); cut float64
    [n, 4] = CutoutAbs' xy with NaN rows for "no Cutout".  Field for field what pack_samples writes (tests compare the bytes)."""
    idx = np.asarray(idx, dtype=np.int64)
    n = idx.shape[0]
    rec = np.zeros(n, dtype=SAMPLE_DTYPE)
    rec["src_index"] = idx
    rec["crop_top"], rec["crop_left"], rec["flip"] = top, left, flip
    rec["cut"][:, 2:] = -1
    if ops is None:
        return rec
    ops = np.asarray(ops, dtype=np.int64)
    val = np.asarray(val, dtype=np.float64)
    k = ops.shape[1]
    if k > 3:
        raise ValueError("srw_aug_sample holds at most 3 ops (RandAugment(3, 5))")
    enh = np.isin(ops, (BRIGHTNESS, COLOR, CONTRAST, SHARPNESS))
    if (val[enh] < 0).any() or (val[ops == SOLARIZE] < 0).any() or (val[ops == SOLARIZE] > 256).any() or (val[ops == POSTERIZE] >= 9).any():
        raise ValueError("op magnitude outside the range the reference asserts (randaugment.py:21-59,114)")
    rec["n_ops"] = k
    o = rec["ops"][:, :k]
    o["op"] = ops
    o["alpha"] = np.where(enh, val, 0.0).astype(np.float32)
    o["ival"] = np.where(ops == POSTERIZE, np.maximum(1, val.astype(np.int64)), np.where(ops == SOLARIZE, np.ceil(val).astype(np.int64), 0))
    a = np.zeros((n, k, 6), dtype=np.float64)
    aff = np.isin(ops, (SHEAR_X, SHEAR_Y, TRANSLATE_X, TRANSLATE_Y))
    a[aff, 0] = 1.0
    a[aff, 4] = 1.0
    a[ops == SHEAR_X, 1] = val[ops == SHEAR_X]
    a[ops == SHEAR_Y, 3] = val[ops == SHEAR_Y]
    a[ops == TRANSLATE_X, 2] = val[ops == TRANSLATE_X] * size
    a[ops == TRANSLATE_Y, 5] = val[ops == TRANSLATE_Y] * size
    ident = np.zeros((n, k), dtype=np.int32)
    for i, j in zip(*np.nonzero(ops == ROTATE)):          # Python floats, as Image.rotate computes them (round(cos, 15) ...)
        m = _affine_coefficients(ROTATE, float(val[i, j]), size)
        if m is None:
            ident[i, j] = 1
        else:
            a[i, j] = m
    geo = np.isin(ops, (ROTATE, SHEAR_X, SHEAR_Y, TRANSLATE_X, TRANSLATE_Y)) & (ident == 0)
    for x, y in ((0, 0), (size, size), (0, size), (size, 0)):   # libImaging's fixed-point range check
        bad = geo & ~((np.abs(x * a[..., 0] + y * a[..., 1] + a[..., 2]) < 32768.0) & (np.abs(x * a[..., 3] + y * a[..., 4] + a[..., 5]) < 32768.0))
        if bad.any():
            raise ValueError("affine coefficients outside the fixed-point range of Image.transform")
    o["a"] = a
    o["identity"] = ident
    if cut is not None:
        cut = np.asarray(cut, dtype=np.float64)
        has = ~np.isnan(cut[:, 0])
        rec["cut"] = np.where(has[:, None], np.nan_to_num(cut).astype(np.int64), np.array([0, 0, -1, -1]))   # ImageDraw truncates the corners
    return rec


def draw_records(indices, size: int, padding: int, strong: bool, rng: np.random.Generator, n_ops: int = 3) -> np.ndarray:
    """Bulk drawing for throughput: the SAME distributions as draw_weak / draw_strong (RandomCrop.get_params, RandomHorizontalFlip,
    RandAugment.__call__, CutoutAbs) from one numpy Generator, vectorised over the batch — not the reference's generator streams, so a
    run is not seed-for-seed the reference's; every record is still transformed exactly as Pillow would transform that decision
    (pack_arrays forms the per-op parameters like pack_samples; Image.rotate's coefficients go through the same Python floats)."""
    idx = np.asarray(indices, dtype=np.int64)
    n = idx.shape[0]
    top = rng.integers(0, 2 * padding + 1, n) if padding > 0 else np.zeros(n, dtype=np.int64)
    left = rng.integers(0, 2 * padding + 1, n) if padding > 0 else np.zeros(n, dtype=np.int64)
    flip = rng.random(n) < 0.5
    if not strong:
        return pack_arrays(idx, top, left, flip, size=size)
    ops = rng.integers(0, 14, (n, n_ops))
    lo = np.array([r[0] for r in _RANGE], dtype=np.float64)[ops]
    hi = np.array([r[1] for r in _RANGE], dtype=np.float64)[ops]
    val = lo + (hi - lo) * rng.random((n, n_ops))
    # Cutout: v = U[0, 0.5) * size; corner = int(max(0, U[0, size) - v / 2)); far corner = min(size, corner + v)
    v = rng.random(n) * 0.5 * size
    x0 = np.maximum(0.0, rng.random(n) * size - v / 2.0).astype(np.int64).astype(np.float64)
    y0 = np.maximum(0.0, rng.random(n) * size - v / 2.0).astype(np.int64).astype(np.float64)
    cut = np.stack([x0, y0, np.minimum(float(size), x0 + v), np.minimum(float(size), y0 + v)], axis=1)
    cut[v <= 0.0] = np.nan
    return pack_arrays(idx, top, left, flip, ops, val, cut, size)


class DeviceImagePipeline:
    """The dataset's uint8 HWC array resident on the device + the two transforms of get_cifar as one kernel launch per batch."""

    def __init__(self, data, mean: Sequence[float], std: Sequence[float], img_size: Optional[int] = None, crop_ratio: float = 0.875,
                 device: str = "cuda"):
        data = torch.as_tensor(np.ascontiguousarray(data)) if not torch.is_tensor(data) else data
        if data.dtype != torch.uint8 or data.ndim != 4 or data.shape[3] != 3 or data.shape[1] != data.shape[2]:
            raise ValueError("expected a uint8 [N, S, S, 3] array (torchvision's CIFAR `.data` layout)")
        self.size = int(data.shape[1])
        if img_size is not None and int(img_size) != self.size:
            raise NotImplementedError("transforms.Resize to a different size (PIL antialiased bilinear) is not built; "
                                      f"source {self.size} px, img_size {img_size}")
        self.padding = int(self.size * (1 - crop_ratio))             # cifar.py:36
        self.lib = L.load()
        self.data = data.contiguous().to(device)
        self.mean = [float(np.float32(m)) for m in mean]
        self.std = [float(np.float32(s)) for s in std]
        self._stage = None                                            # pinned staging buffer of the decision records

    def _upload(self, arr) -> torch.Tensor:
        nbytes = C.sizeof(arr)
        if self._stage is None or self._stage.numel() < nbytes:
            self._stage = torch.empty(max(nbytes, 64 * C.sizeof(L.AugSample)), dtype=torch.uint8).pin_memory()
        host = self._stage[:nbytes]
        host.copy_(torch.frombuffer(memoryview(arr).cast("B"), dtype=torch.uint8))
        dev = torch.empty(nbytes, dtype=torch.uint8, device=self.data.device)
        dev.copy_(host, non_blocking=True)
        self._stage_event = torch.cuda.Event()
        self._stage_event.record()
        return dev

    def transform(self, indices: Sequence[int], decisions: Sequence[AugDecision], out: Optional[torch.Tensor] = None,
                  return_u8: bool = False):
        idx = [int(i) for i in indices]
        if not idx:
            raise ValueError("empty batch")
        if min(idx) < 0 or max(idx) >= self.data.shape[0]:
            raise IndexError("sample index outside the dataset")
        for d in decisions:
            if not (0 <= d.crop_top <= 2 * self.padding and 0 <= d.crop_left <= 2 * self.padding):
                raise ValueError("crop offset outside the padded image")
        n, S = len(idx), self.size
        if getattr(self, "_stage_event", None) is not None:
            self._stage_event.synchronize()                           # the previous batch's records have left the staging buffer
        recs = self._upload(pack_samples(idx, decisions, S))
        if out is None:
            out = torch.empty(n, 3, S, S, dtype=torch.float32, device=self.data.device)
        assert out.is_contiguous() and out.shape == (n, 3, S, S) and out.dtype == torch.float32
        u8 = torch.empty(n, S, S, 3, dtype=torch.uint8, device=self.data.device) if return_u8 else None
        a = L.AugmentArgs(src=L.ptr(self.data), n_src=self.data.shape[0], img_size=S, padding=self.padding, samples=L.ptr(recs), n=n,
                          mean=(L.f32 * 3)(*self.mean), std=(L.f32 * 3)(*self.std), out=L.ptr(out), out_u8=L.ptr(u8))
        L.check(self.lib.srw_augment_batch(C.byref(a), L.stream_ptr()), "srw_augment_batch")
        recs.record_stream(torch.cuda.current_stream())
        return (out, u8) if return_u8 else out

    def transform_records(self, rec: np.ndarray, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Same launch as transform() from a SAMPLE_DTYPE array (records_from_decisions / draw_records)."""
        if rec.dtype != SAMPLE_DTYPE or rec.ndim != 1 or rec.shape[0] == 0:
            raise ValueError("expected a non-empty 1-D SAMPLE_DTYPE array")
        if rec["src_index"].min() < 0 or rec["src_index"].max() >= self.data.shape[0]:
            raise IndexError("sample index outside the dataset")
        if rec["crop_top"].min() < 0 or rec["crop_top"].max() > 2 * self.padding or rec["crop_left"].min() < 0 or rec["crop_left"].max() > 2 * self.padding:
            raise ValueError("crop offset outside the padded image")
        n, S = int(rec.shape[0]), self.size
        raw = torch.from_numpy(np.ascontiguousarray(rec).view(np.uint8))
        if getattr(self, "_stage_event", None) is not None:
            self._stage_event.synchronize()
        if self._stage is None or self._stage.numel() < raw.numel():
            self._stage = torch.empty(max(raw.numel(), 64 * SAMPLE_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
        host = self._stage[:raw.numel()]
        host.copy_(raw)
        recs = torch.empty(raw.numel(), dtype=torch.uint8, device=self.data.device)
        recs.copy_(host, non_blocking=True)
        self._stage_event = torch.cuda.Event()
        self._stage_event.record()
        if out is None:
            out = torch.empty(n, 3, S, S, dtype=torch.float32, device=self.data.device)
        assert out.is_contiguous() and out.shape == (n, 3, S, S) and out.dtype == torch.float32
        a = L.AugmentArgs(src=L.ptr(self.data), n_src=self.data.shape[0], img_size=S, padding=self.padding, samples=L.ptr(recs), n=n,
                          mean=(L.f32 * 3)(*self.mean), std=(L.f32 * 3)(*self.std), out=L.ptr(out), out_u8=None)
        L.check(self.lib.srw_augment_batch(C.byref(a), L.stream_ptr()), "srw_augment_batch")
        recs.record_stream(torch.cuda.current_stream())
        return out

    def weak_and_strong_fast(self, indices: Sequence[int], rng: np.random.Generator) -> Tuple[torch.Tensor, torch.Tensor]:
        """Both views of every sample in one launch, decisions drawn in bulk (draw_records): the throughput route."""
        idx = np.asarray(indices, dtype=np.int64)
        rec = np.concatenate([draw_records(idx, self.size, self.padding, False, rng), draw_records(idx, self.size, self.padding, True, rng)])
        both = self.transform_records(rec)
        return both[:idx.shape[0]], both[idx.shape[0]:]

    def val(self, indices: Sequence[int]) -> torch.Tensor:
        """transform_val (cifar.py:51-55): Resize (identity) -> ToTensor -> Normalize; nothing random."""
        return self.transform(indices, [AugDecision(self.padding, self.padding, False)] * len(indices))

    # the reference draws per sample, in __getitem__ order
    def weak(self, indices: Sequence[int]) -> torch.Tensor:
        return self.transform(indices, [draw_weak(self.size, self.padding) for _ in indices])

    def weak_and_strong(self, indices: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
        """datasetbase.py:90,115: per sample, transform(img) then strong_transform(img); both views in ONE launch."""
        dec_w, dec_s = [], []
        for _ in indices:
            dec_w.append(draw_weak(self.size, self.padding))
            dec_s.append(draw_strong(self.size, self.padding))
        n = len(dec_w)
        both = self.transform(list(indices) + list(indices), dec_w + dec_s)
        return both[:n], both[n:]


class DeviceSSLLoader:
    """What `zip(loader_dict['train_lb'], loader_dict['train_ulb'])` yields in AlgorithmBase.train() (core/algorithmbase.py:357-370),
    with the batches already on the device: ({'idx_lb', 'x_lb', 'y_lb'}, {'idx_ulb', 'x_ulb_w', 'x_ulb_s'}).  Index streams come from
    the caller's samplers (any iterable of index lists, e.g. torch BatchSampler over the reference's DistributedSampler)."""

    def __init__(self, lb: DeviceImagePipeline, lb_targets, ulb: DeviceImagePipeline, lb_batches: Iterable[Sequence[int]],
                 ulb_batches: Iterable[Sequence[int]], bulk_rng: Optional[np.random.Generator] = None):
        """bulk_rng: draw the decisions of a whole batch from this numpy Generator (draw_records: same distributions, 20x the host
        throughput) instead of per sample from the generators the reference's transforms consume."""
        self.lb, self.ulb = lb, ulb
        self.targets = torch.as_tensor(np.asarray(lb_targets), dtype=torch.int64).to(lb.data.device)
        self.lb_batches, self.ulb_batches = lb_batches, ulb_batches
        self.bulk_rng = bulk_rng

    def __iter__(self) -> Iterator[Tuple[dict, dict]]:
        for ib, iu in zip(self.lb_batches, self.ulb_batches):
            ib_t = torch.as_tensor(list(ib), dtype=torch.int64)
            iu_t = torch.as_tensor(list(iu), dtype=torch.int64)
            if self.bulk_rng is None:
                x_lb = self.lb.weak(ib)                               # the labelled loader's batch is collated first
                x_w, x_s = self.ulb.weak_and_strong(iu)
            else:
                x_lb = self.lb.transform_records(draw_records(list(ib), self.lb.size, self.lb.padding, False, self.bulk_rng))
                x_w, x_s = self.ulb.weak_and_strong_fast(list(iu), self.bulk_rng)
            dev = self.lb.data.device
            yield ({"idx_lb": ib_t.to(dev), "x_lb": x_lb, "y_lb": self.targets[ib_t.to(dev)]},
                   {"idx_ulb": iu_t.to(dev), "x_ulb_w": x_w, "x_ulb_s": x_s})
