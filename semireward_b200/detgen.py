"""Deterministic, platform-independent synthetic tensors.

Golden fixtures are produced in the build container from the live reference and checked on the
GPU box, where neither the reference nor its RNG stream exists.  Everything that feeds a parity
test (weights, images, labels, indices) therefore comes from this counter-based generator: pure
integer hashing (splitmix64) followed by exact float64 arithmetic, so the same (tag, shape) gives
bit-identical float32 tensors on every machine and numpy version.

Values are the sum of four uniforms rescaled to unit variance (Irwin-Hall, |v| <= 2*sqrt(3)):
no transcendental function is involved, hence no libm dependence.
"""
from __future__ import annotations

import zlib

import numpy as np

_SQRT3 = 1.7320508075688772
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        z = z ^ (z >> np.uint64(31))
    return z


def _tag_seed(tag: str, seed: int) -> np.uint64:
    h = zlib.crc32(tag.encode("utf-8")) & 0xFFFFFFFF
    return np.uint64(((seed & 0xFFFFFFFF) << 32) | h)


def _uniform53(tag: str, seed: int, n: int, stream: int) -> np.ndarray:
    """n float64 uniforms in [0,1) for (tag, seed, stream)."""
    base = _splitmix64(np.array([_tag_seed(tag, seed)], dtype=np.uint64) + np.uint64(stream * 0x51ED27))[0]
    with np.errstate(over="ignore"):
        ctr = (np.arange(n, dtype=np.uint64) * np.uint64(0xD1342543DE82EF95) + base) & _M64
    z = _splitmix64(ctr)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def normal(tag: str, shape, seed: int = 0, std: float = 1.0, mean: float = 0.0) -> np.ndarray:
    """float32 array ~ unit-variance bell shape (Irwin-Hall n=4), scaled by std, shifted by mean."""
    n = int(np.prod(shape)) if len(tuple(shape)) else 1
    acc = np.zeros(n, dtype=np.float64)
    for s in range(4):
        acc += _uniform53(tag, seed, n, s)
    v = (acc - 2.0) * _SQRT3
    return (v * std + mean).astype(np.float32).reshape(shape)


def uniform(tag: str, shape, seed: int = 0, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    n = int(np.prod(shape)) if len(tuple(shape)) else 1
    u = _uniform53(tag, seed, n, 7)
    return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


def integers(tag: str, shape, lo: int, hi: int, seed: int = 0) -> np.ndarray:
    """int64 array uniform in [lo, hi)."""
    n = int(np.prod(shape)) if len(tuple(shape)) else 1
    u = _uniform53(tag, seed, n, 11)
    return (lo + np.floor(u * (hi - lo)).astype(np.int64)).reshape(shape)


def distinct_integers(tag: str, n: int, hi: int, seed: int = 0) -> np.ndarray:
    """n distinct int64 draws from [0, hi) (rank of hashed keys -> a permutation prefix)."""
    assert n <= hi
    u = _uniform53(tag, seed, hi, 13)
    return np.argsort(u, kind="stable")[:n].astype(np.int64)


# --------------------------------------------------------------------------------------------
# Parameter fills.  Rules are keyed on the reference's state_dict names
# (semilearn/nets/vit/vit.py:232-275; semilearn/algorithms/semireward/semireward.py:30-50, 9-19).
# Scales are chosen so that activations stay O(1) through 12 pre-LN blocks and logits have a
# spread of a few units (so softmax/argmax/thresholds are exercised away from ties).
# --------------------------------------------------------------------------------------------

def fill_param(name: str, shape, seed: int = 0) -> np.ndarray:
    shape = tuple(int(s) for s in shape)
    if name.endswith("cls_token") or name.endswith("pos_embed"):
        return normal(name, shape, seed, std=0.02)
    is_bn = name.endswith((".bn1.weight", ".bn2.weight")) or name == "bn1.weight"   # WideResNet's BatchNorm scales (wrn.py:33,37,97)
    if ("norm" in name or "LayerNorm" in name or is_bn) and name.endswith(".weight") and len(shape) == 1:   # "LayerNorm": the HF BERT names (bert.py:13)
        return normal(name, shape, seed, std=0.1, mean=1.0)
    if name.endswith(".bias") or len(shape) == 1:
        return normal(name, shape, seed, std=0.05)
    if "label_embedding" in name:
        return normal(name, shape, seed, std=1.0)
    # weights: fan_in = prod(shape[1:])
    fan_in = int(np.prod(shape[1:]))
    gain = 2.0 if name.startswith("head") else 1.0
    return normal(name, shape, seed, std=gain / np.sqrt(fan_in))


def fill_state_dict(named_shapes, seed: int = 0, prefix: str = ""):
    """{name: float32 ndarray} for an iterable of (name, shape)."""
    return {n: fill_param(prefix + n, s, seed) for n, s in named_shapes}


def ssl_batch(batch_size: int, uratio: int, num_classes: int, ulb_dest_len: int, img_size: int = 32,
              in_chans: int = 3, seed: int = 1, step: int = 0):
    """One synthetic SSL batch with the reference's batch-dict keys
    (semilearn/datasets/cv_datasets/datasetbase.py:90-111)."""
    bu = batch_size * uratio
    s = seed * 1000003 + step
    return dict(
        x_lb=normal("x_lb", (batch_size, in_chans, img_size, img_size), s),
        y_lb=integers("y_lb", (batch_size,), 0, num_classes, s),
        idx_ulb=distinct_integers("idx_ulb", bu, ulb_dest_len, s),
        x_ulb_w=normal("x_ulb_w", (bu, in_chans, img_size, img_size), s),
        x_ulb_s=normal("x_ulb_s", (bu, in_chans, img_size, img_size), s),
    )


def nlp_batch(batch_size: int, uratio: int, num_classes: int, ulb_dest_len: int, max_length: int = 512, vocab_size: int = 30522,
              min_length: int = 0, seed: int = 1, step: int = 0):
    """One synthetic text SSL batch with the reference's batch-dict keys (semilearn/datasets/nlp_datasets/datasetbase.py,
    collators: x_* = {'input_ids', 'attention_mask'} int64 [B, L]): token ids uniform in [1000, vocab) (or [1, vocab) for a
    small test vocabulary), lengths uniform in [min_length or L/4, L], a zero tail of padding (pad id 0) under a zero
    attention mask — BASELINE configs[3] as SURVEY.md §8d describes it."""
    bu = batch_size * uratio
    s = seed * 1000003 + step
    lo = 1000 if vocab_size > 2000 else 1
    min_len = min_length if min_length > 0 else max(1, max_length // 4)

    def text(tag, n):
        ids = integers(tag + "_ids", (n, max_length), lo, vocab_size, s)
        lens = integers(tag + "_len", (n,), min_len, max_length + 1, s)
        am = (np.arange(max_length, dtype=np.int64)[None, :] < lens[:, None]).astype(np.int64)
        return dict(input_ids=ids * am, attention_mask=am)
    return dict(
        x_lb=text("x_lb", batch_size),
        y_lb=integers("y_lb", (batch_size,), 0, num_classes, s),
        idx_ulb=distinct_integers("idx_ulb", bu, ulb_dest_len, s),
        x_ulb_w=text("x_ulb_w", bu),
        x_ulb_s=text("x_ulb_s", bu),
    )


def audio_batch(batch_size: int, uratio: int, num_classes: int, ulb_dest_len: int, samples: int = 64000, seed: int = 1, step: int = 0):
    """One synthetic audio SSL batch with the reference's batch-dict keys (semilearn/datasets/audio_datasets/datasetbase.py:
    x_* = fp32 waveforms [B, samples], 16 kHz): unit-variance noise, the normalised case of BASELINE configs[4]."""
    bu = batch_size * uratio
    s = seed * 1000003 + step
    return dict(
        x_lb=normal("x_lb", (batch_size, samples), s),
        y_lb=integers("y_lb", (batch_size,), 0, num_classes, s),
        idx_ulb=distinct_integers("idx_ulb", bu, ulb_dest_len, s),
        x_ulb_w=normal("x_ulb_w", (bu, samples), s),
        x_ulb_s=normal("x_ulb_s", (bu, samples), s),
    )
