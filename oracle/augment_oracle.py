"""TEST INFRASTRUCTURE — CPU restatement (numpy, integer / byte arithmetic) of the reference's input pipeline for the image
configs: `transform_weak` / `transform_strong` of `get_cifar`
(semilearn/datasets/cv_datasets/cifar.py:34-49) as called by `BasicDataset.__getitem__`
(semilearn/datasets/cv_datasets/datasetbase.py:74-115):

    Resize(img_size)  [identity when the source already has that size]  ->  RandomCrop(img_size, padding, 'reflect')
    -> RandomHorizontalFlip -> [strong only: RandAugment(3, 5) = 3 ops of `augment_list()` + Cutout
    (semilearn/datasets/augmentation/randaugment.py:16-206)] -> ToTensor -> Normalize(mean, std)

The arithmetic behind the reference's calls lives in third-party code that is not under /root/reference: Pillow
(`requirements.txt` unpinned; 12.2.0 installed here = the de-facto pin) and torchvision 0.26.  Their published algorithms are
restated below (ImageOps.autocontrast / equalize / posterize / solarize look-up tables, ImageEnhance = Image.blend in C
float, RGB->L `(19595 R + 38470 G + 7471 B + 0x8000) >> 16`, ImageFilter.SMOOTH, Image.rotate's matrix, the nearest-
neighbour affine transform in 16.16 fixed point / the axis-aligned "scale" path, ImageDraw.rectangle) and PINNED against
the live libraries in this container: tests/test_augment_oracle.py (every op over random and degenerate images and the
reference's own `RandAugment.__call__` / transforms under seeded RNGs), fixture tests/golden/augment_cifar.npz made by
tests/golden/make_golden_augment.py from the imported reference.

Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module; the product never does.
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# op ids = position in the reference's augment_list() (randaugment.py:157-174)
AUTOCONTRAST, BRIGHTNESS, COLOR, CONTRAST, EQUALIZE, IDENTITY, POSTERIZE, ROTATE, SHARPNESS, SHEAR_X, SHEAR_Y, SOLARIZE, \
    TRANSLATE_X, TRANSLATE_Y = range(14)
OP_NAMES = ["AutoContrast", "Brightness", "Color", "Contrast", "Equalize", "Identity", "Posterize", "Rotate", "Sharpness",
            "ShearX", "ShearY", "Solarize", "TranslateX", "TranslateY"]
OP_RANGE = [(0, 1), (0.05, 0.95), (0.05, 0.95), (0.05, 0.95), (0, 1), (0, 1), (4, 8), (-30, 30), (0.05, 0.95), (-0.3, 0.3),
            (-0.3, 0.3), (0, 256), (-0.3, 0.3), (-0.3, 0.3)]           # (min_val, max_val) of augment_list()
CUTOUT_COLOR = (125, 123, 114)                                          # randaugment.py:148

CIFAR100_MEAN = [x / 255 for x in [129.3, 124.1, 112.4]]               # cifar.py:18-21
CIFAR100_STD = [x / 255 for x in [68.2, 65.4, 70.4]]
CIFAR10_MEAN, CIFAR10_STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


@dataclass
class Decision:
    """Everything random about one transformed sample (what the product's kernel takes as explicit input)."""
    crop_top: int = 0
    crop_left: int = 0
    flip: bool = False
    ops: List[Tuple[int, float]] = field(default_factory=list)         # (op id, val) in application order
    cutout: Optional[Tuple[float, float, float, float]] = None         # xy of CutoutAbs (floats, as handed to ImageDraw)


# ------------------------------------------------------------------------------------------------------------------------------
# decisions, drawn from the same generators in the same order as the reference
# ------------------------------------------------------------------------------------------------------------------------------
def draw_crop_flip(size: int, padding: int) -> Tuple[int, int, bool]:
    """torchvision RandomCrop.get_params + RandomHorizontalFlip.forward: two torch.randint draws (none when nothing to choose),
    then torch.rand(1) < 0.5."""
    import torch
    span = 2 * padding
    if span == 0:
        top = left = 0
    else:
        top = int(torch.randint(0, span + 1, size=(1,)).item())
        left = int(torch.randint(0, span + 1, size=(1,)).item())
    flip = bool(torch.rand(1) < 0.5)
    return top, left, flip


def draw_randaugment(size: int, n: int = 3) -> Tuple[List[Tuple[int, float]], Optional[Tuple[float, float, float, float]]]:
    """RandAugment.__call__ (randaugment.py:196-203): `random.choices` for the ops, one `random.random()` per op, one for the Cutout
    size, then CutoutAbs' two `np.random.uniform` draws (randaugment.py:136-146; none when the size is 0)."""
    ids = random.choices(range(14), k=n)
    ops = []
    for i in ids:
        lo, hi = OP_RANGE[i]
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
    return ops, cut


def draw_weak(size: int, padding: int) -> Decision:
    t, l, f = draw_crop_flip(size, padding)
    return Decision(t, l, f)


def draw_strong(size: int, padding: int, n: int = 3) -> Decision:
    t, l, f = draw_crop_flip(size, padding)
    ops, cut = draw_randaugment(size, n)
    return Decision(t, l, f, ops, cut)


# ------------------------------------------------------------------------------------------------------------------------------
# geometry front: reflect padding, crop, flip
# ------------------------------------------------------------------------------------------------------------------------------
def _reflect(i: np.ndarray, n: int) -> np.ndarray:
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def crop_flip(img: np.ndarray, top: int, left: int, flip: bool, padding: int) -> np.ndarray:
    """F.pad(img, padding, padding_mode='reflect') (numpy 'reflect': the edge pixel is not repeated), crop at (top, left), hflip."""
    h, w, _ = img.shape
    ys = _reflect(np.arange(h) + top - padding, h)
    xs = np.arange(w)
    if flip:
        xs = w - 1 - xs
    xs = _reflect(xs + left - padding, w)
    return img[ys][:, xs]


# ------------------------------------------------------------------------------------------------------------------------------
# look-up-table ops (PIL.ImageOps)
# ------------------------------------------------------------------------------------------------------------------------------
def _hist(img: np.ndarray) -> np.ndarray:
    return np.stack([np.bincount(img[..., c].ravel(), minlength=256) for c in range(3)])


def autocontrast_lut(h: Sequence[int]) -> np.ndarray:
    """ImageOps.autocontrast(cutoff=0), one band: stretch [lowest, highest] occupied level to [0, 255] in Python floats."""
    lo = next(i for i in range(256) if h[i])
    hi = next(i for i in range(255, -1, -1) if h[i])
    if hi <= lo:
        return np.arange(256, dtype=np.uint8)
    scale = 255.0 / (hi - lo)
    offset = -lo * scale
    return np.array([min(255, max(0, int(ix * scale + offset))) for ix in range(256)], dtype=np.uint8)


def equalize_lut(h: Sequence[int]) -> np.ndarray:
    """ImageOps.equalize, one band (integer arithmetic throughout)."""
    nz = [int(v) for v in h if v]
    if len(nz) <= 1:
        return np.arange(256, dtype=np.uint8)
    step = (sum(nz) - nz[-1]) // 255
    if not step:
        return np.arange(256, dtype=np.uint8)
    lut, n = [], step // 2
    for i in range(256):
        lut.append(n // step)
        n += int(h[i])
    return np.minimum(np.array(lut, dtype=np.int64), 255).astype(np.uint8)   # entries can pass 255 (step is floored); Image.point clips them


def _apply_luts(img: np.ndarray, luts: Sequence[np.ndarray]) -> np.ndarray:
    return np.stack([luts[c][img[..., c]] for c in range(3)], axis=-1)


def autocontrast(img):
    h = _hist(img)
    return _apply_luts(img, [autocontrast_lut(h[c]) for c in range(3)])


def equalize(img):
    h = _hist(img)
    return _apply_luts(img, [equalize_lut(h[c]) for c in range(3)])


def posterize(img, v: float):
    bits = max(1, int(v))
    return img & np.uint8(~(2 ** (8 - bits) - 1) & 0xFF)


def solarize(img, v: float):
    i = img.astype(np.int64)
    return np.where(i < v, i, 255 - i).astype(np.uint8)


# ------------------------------------------------------------------------------------------------------------------------------
# ImageEnhance = Image.blend(degenerate, image, factor) in C float
# ------------------------------------------------------------------------------------------------------------------------------
def blend(deg: np.ndarray, img: np.ndarray, factor: float) -> np.ndarray:
    """libImaging Blend.c: out = (UINT8)((int)in1 + alpha * ((int)in2 - (int)in1)) with a C `float` alpha (interpolating branch;
    clipped when alpha is outside [0, 1]).  float32 product, float32 sum, truncation."""
    alpha = np.float32(factor)
    d = (img.astype(np.int32) - deg.astype(np.int32)).astype(np.float32)
    t = deg.astype(np.float32) + alpha * d
    if 0.0 <= factor <= 1.0:
        return t.astype(np.int32).astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int32))).astype(np.uint8)


def luma(img: np.ndarray) -> np.ndarray:
    """convert('L'): ITU-R 601-2 in 16-bit fixed point with rounding."""
    i = img.astype(np.int64)
    return ((i[..., 0] * 19595 + i[..., 1] * 38470 + i[..., 2] * 7471 + 0x8000) >> 16).astype(np.uint8)


def brightness(img, v):
    return blend(np.zeros_like(img), img, v)


def color(img, v):
    g = luma(img)
    return blend(np.repeat(g[..., None], 3, axis=-1), img, v)


def contrast(img, v):
    g = luma(img)
    mean = int(int(g.astype(np.int64).sum()) / g.size + 0.5)          # ImageStat mean (sum / count in Python floats) + 0.5, truncated
    return blend(np.full_like(img, mean), img, v)


def smooth(img: np.ndarray) -> np.ndarray:
    """ImageFilter.SMOOTH = 3x3 kernel (1 1 1 / 1 5 1 / 1 1 1) / 13; the one-pixel border is copied.  Filter.c rounds the float sum
    (+0.5, truncate); sum/13 is never within float error of a half, so round(sum/13) = (2 sum + 13) // 26 exactly."""
    h, w, _ = img.shape
    out = img.copy()
    i = img.astype(np.int64)
    acc = np.zeros((h - 2, w - 2, 3), dtype=np.int64)
    for dy in range(3):
        for dx in range(3):
            acc += (5 if dy == 1 and dx == 1 else 1) * i[dy:dy + h - 2, dx:dx + w - 2]
    out[1:-1, 1:-1] = (2 * acc + 13) // 26
    return out


def sharpness(img, v):
    return blend(smooth(img), img, v)


# ------------------------------------------------------------------------------------------------------------------------------
# affine ops: Image.rotate / Image.transform(AFFINE), nearest neighbour, fill 0
# ------------------------------------------------------------------------------------------------------------------------------
def rotate_matrix(angle: float, w: int, h: int) -> Optional[List[float]]:
    """Image.rotate (no expand, centre = (w/2, h/2)): the inverse map's coefficients in Python floats, rounded to 15 decimals as
    Pillow does.  None = the angle is a multiple of 360 (Image.rotate returns a copy)."""
    angle = angle % 360.0
    if angle == 0:
        return None
    if angle in (90, 180, 270):
        raise NotImplementedError("transpose fast paths of Image.rotate: val is a continuous draw, never hit")
    cx, cy = w / 2.0, h / 2.0
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    m[2] = m[0] * (-cx) + m[1] * (-cy) + m[2]
    m[5] = m[3] * (-cx) + m[4] * (-cy) + m[5]
    m[2] += cx
    m[5] += cy
    return m


def op_matrix(op: int, v: float, w: int, h: int) -> Optional[List[float]]:
    if op == ROTATE:
        return rotate_matrix(v, w, h)
    if op == SHEAR_X:
        return [1, v, 0, 0, 1, 0]
    if op == SHEAR_Y:
        return [1, 0, 0, v, 1, 0]
    if op == TRANSLATE_X:
        return [1, 0, v * w, 0, 1, 0]
    if op == TRANSLATE_Y:
        return [1, 0, 0, 0, 1, v * h]
    raise ValueError(op)


def affine_nearest(img: np.ndarray, a: Sequence[float]) -> np.ndarray:
    """libImaging Geometry.c, ImagingTransformAffine with the NEAREST filter.
    a[1] == a[3] == 0 -> ImagingScaleAffine: source column / row = trunc of a coordinate accumulated in doubles (negative -> outside).
    otherwise -> affine_fixed: 16.16 fixed point, coefficients FLOOR(v * 65536 + 0.5), pixel centre folded into the offsets."""
    h, w, _ = img.shape
    out = np.zeros_like(img)
    a = [float(v) for v in a]
    if a[1] == 0 and a[3] == 0:
        xo = a[2] + a[0] * 0.5
        xin = []
        for _ in range(w):
            xin.append(-1 if xo < 0.0 else int(xo))
            xo += a[0]
        yo = a[5] + a[4] * 0.5
        for y in range(h):
            yi = -1 if yo < 0.0 else int(yo)
            if 0 <= yi < h:
                for x in range(w):
                    if 0 <= xin[x] < w:
                        out[y, x] = img[yi, xin[x]]
            yo += a[4]
        return out

    def fix(v):
        return int(math.floor(v * 65536.0 + 0.5))
    a0, a1, a3, a4 = fix(a[0]), fix(a[1]), fix(a[3]), fix(a[4])
    a2 = fix(a[2] + a[0] * 0.5 + a[1] * 0.5)
    a5 = fix(a[5] + a[3] * 0.5 + a[4] * 0.5)
    ys, xs = np.mgrid[0:h, 0:w]
    xx = (a2 + a1 * ys + a0 * xs) >> 16
    yy = (a5 + a4 * ys + a3 * xs) >> 16
    ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
    out[ok] = img[yy[ok], xx[ok]]
    return out


def cutout_abs(img: np.ndarray, xy: Sequence[float]) -> np.ndarray:
    """ImageDraw.rectangle(xy, fill): corner coordinates truncated to int, both corners INCLUSIVE, clipped to the image."""
    h, w, _ = img.shape
    x0, y0, x1, y1 = (int(v) for v in xy)
    out = img.copy()
    if x1 < x0 or y1 < y0:
        return out
    out[max(y0, 0):min(y1, h - 1) + 1, max(x0, 0):min(x1, w - 1) + 1] = np.array(CUTOUT_COLOR, dtype=np.uint8)
    return out


def apply_op(img: np.ndarray, op: int, v: float) -> np.ndarray:
    h, w, _ = img.shape
    if op == AUTOCONTRAST:
        return autocontrast(img)
    if op == BRIGHTNESS:
        return brightness(img, v)
    if op == COLOR:
        return color(img, v)
    if op == CONTRAST:
        return contrast(img, v)
    if op == EQUALIZE:
        return equalize(img)
    if op == IDENTITY:
        return img
    if op == POSTERIZE:
        return posterize(img, v)
    if op == SHARPNESS:
        return sharpness(img, v)
    if op == SOLARIZE:
        return solarize(img, v)
    m = op_matrix(op, v, w, h)
    return img.copy() if m is None else affine_nearest(img, m)


# ------------------------------------------------------------------------------------------------------------------------------
# whole pipeline
# ------------------------------------------------------------------------------------------------------------------------------
def transform_u8(img: np.ndarray, d: Decision, padding: int) -> np.ndarray:
    """uint8 HWC in -> uint8 HWC out: everything up to (not including) ToTensor."""
    out = crop_flip(img, d.crop_top, d.crop_left, d.flip, padding)
    for op, v in d.ops:
        out = apply_op(out, op, v)
    if d.cutout is not None:
        out = cutout_abs(out, d.cutout)
    return out


def to_tensor_normalize(img: np.ndarray, mean: Sequence[float], std: Sequence[float]) -> np.ndarray:
    """ToTensor (uint8 HWC -> float32 CHW, true division by 255) then Normalize: (t - mean) / std with float32 mean / std."""
    t = img.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    m = np.asarray(mean, dtype=np.float32)[:, None, None]
    s = np.asarray(std, dtype=np.float32)[:, None, None]
    return ((t - m) / s).astype(np.float32)


def transform(img: np.ndarray, d: Decision, padding: int, mean: Sequence[float], std: Sequence[float]) -> np.ndarray:
    return to_tensor_normalize(transform_u8(img, d, padding), mean, std)
