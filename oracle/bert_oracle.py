"""TEST INFRASTRUCTURE — CPU restatement of the text path of the hot loop (BASELINE configs[3]):
`ClassificationBert` (semilearn/nets/bert/bert.py:9-48) around a BERT encoder, driven by the same SSL step as the vision
path but with `use_cat: False` (three separate backbone calls, srsoftmatch.py:119-130 and twins).

Only tests/ (and, later, smoke()/bench.py's CPU arm) may import this module; the product never does.

Third-party arithmetic: the encoder is Hugging Face `transformers.BertModel` — `transformers>=4.30.0` in
the reference's requirements.txt (unpinned), 5.5.0 installed in this image, which is the de-facto pin.  It is not under
/root/reference, so its published algorithm (Devlin et al. 2018; modeling_bert.py of the installed version, eager
attention) is restated here in plain torch ops in the same order, and pinned against the live `ClassificationBert`
running on CPU in this container (tests/test_bert_oracle.py, fixtures tests/golden/bert_*.npz made by
tests/golden/make_golden_bert.py).  Deterministic parity mode = all dropout probabilities 0 (the reference trains with
hidden / attention-probs / pooled dropout 0.1; like DropPath for the ViT, a stochastic pass can only be compared
statistically or with injected masks: `BertDropout` below fixes the order in which masks are drawn).

Layout of one forward (B sequences of L tokens, hidden 768, 12 heads of 64):
  e   = word[ids] + type[0] + pos[:L]            -> LN(eps 1e-12) -> drop          (BertEmbeddings)
  per layer (post-LN):  a = drop(softmax(q k^T / 8 + keymask)) v ; x = LN(drop(a Wo + bo) + x)
                        x = LN(drop(gelu(x W1 + b1) W2 + b2) + x)
  feat   = mean over ALL L positions (padding included, bert.py:36-37) of drop(x)
  logits = gelu(feat Wc1 + bc1) Wc2 + bc2                                           (bert.py:16-20)
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import ssl_oracle as O

Tensor = torch.Tensor


@dataclass
class BertCfg:
    vocab_size: int = 30522
    hidden: int = 768                 # ClassificationBert hard-codes 768 (bert.py:15-19)
    layers: int = 12
    heads: int = 12
    intermediate: int = 3072
    max_position: int = 512
    type_vocab: int = 2
    eps: float = 1e-12
    hidden_dropout: float = 0.1       # BertConfig.hidden_dropout_prob
    attn_dropout: float = 0.1         # BertConfig.attention_probs_dropout_prob
    pooled_dropout: float = 0.1       # bert.py:14
    num_classes: int = 2

    def param_shapes(self) -> List[Tuple[str, Tuple[int, ...]]]:
        """`ClassificationBert.state_dict()` parameter order (HF BertModel with pooler, then the classifier)."""
        H, I = self.hidden, self.intermediate
        out = [("bert.embeddings.word_embeddings.weight", (self.vocab_size, H)),
               ("bert.embeddings.position_embeddings.weight", (self.max_position, H)),
               ("bert.embeddings.token_type_embeddings.weight", (self.type_vocab, H)),
               ("bert.embeddings.LayerNorm.weight", (H,)), ("bert.embeddings.LayerNorm.bias", (H,))]
        for i in range(self.layers):
            p = f"bert.encoder.layer.{i}."
            out += [(p + "attention.self.query.weight", (H, H)), (p + "attention.self.query.bias", (H,)),
                    (p + "attention.self.key.weight", (H, H)), (p + "attention.self.key.bias", (H,)),
                    (p + "attention.self.value.weight", (H, H)), (p + "attention.self.value.bias", (H,)),
                    (p + "attention.output.dense.weight", (H, H)), (p + "attention.output.dense.bias", (H,)),
                    (p + "attention.output.LayerNorm.weight", (H,)), (p + "attention.output.LayerNorm.bias", (H,)),
                    (p + "intermediate.dense.weight", (I, H)), (p + "intermediate.dense.bias", (I,)),
                    (p + "output.dense.weight", (H, I)), (p + "output.dense.bias", (H,)),
                    (p + "output.LayerNorm.weight", (H,)), (p + "output.LayerNorm.bias", (H,))]
        out += [("bert.pooler.dense.weight", (H, H)), ("bert.pooler.dense.bias", (H,)),   # built by BertModel, unused by bert.py:34-37
                ("classifier.0.weight", (H, H)), ("classifier.0.bias", (H,)),
                ("classifier.2.weight", (self.num_classes, H)), ("classifier.2.bias", (self.num_classes,))]
        return out

    def fwd_flops_per_seq(self, L: int) -> float:
        """2*MACs of one sequence: linear layers + the two attention products (96.64 GF at L = 512)."""
        H, I = self.hidden, self.intermediate
        per_layer = 2.0 * L * (4 * H * H + 2 * H * I) + 4.0 * L * L * H
        return self.layers * per_layer + 2.0 * (H * H + H * self.num_classes)


def lowbias32(x):
    """32-bit integer mixer (two multiply-xorshift rounds) on numpy uint32 arrays / ints; wraps modulo 2^32."""
    import numpy as np
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def call_key(seed: int, call: int) -> int:
    """Stream key of one backbone call (the reference makes three per step with `use_cat: False`, and three more per sampling
    pass): every call draws from its own stream, like successive nn.Dropout invocations do."""
    return int(lowbias32((seed * 0x9E3779B9 + call * 0x85EBCA6B + 0x165667B1) & 0xFFFFFFFF))


def counter_keep_mask(numel: int, keep: float, key: int, site: int, offset: int = 0) -> Tensor:
    """Counter-based Bernoulli(keep) mask that a CUDA epilogue regenerates element for element (no RNG stream to share; the
    definition is restated from include/srw.h `srw_dropout`): element i of dropout site `site` of the stream `key` is kept iff
    the top 24 bits of lowbias32(i + lowbias32(key + site * 0x9E3779B9)) are below keep * 2^24.  Pure 32-bit integer arithmetic."""
    import numpy as np
    site_key = int(lowbias32((key + site * 0x9E3779B9) & 0xFFFFFFFF))
    idx = (np.arange(offset, offset + numel, dtype=np.uint64) + site_key) & 0xFFFFFFFF
    return torch.from_numpy((lowbias32(idx) >> 8) < int(keep * (1 << 24)))


class BertDropout:
    """Order in which a stochastic pass draws its Bernoulli(1-p)/(1-p) masks: embeddings; per layer attention probabilities
    [B, heads, L, L], attention output [B, L, H], FFN output [B, L, H]; last the pooled-feature dropout [B, L, H].
    p = 0 everywhere (deterministic parity mode) draws nothing.  `generator` is either a torch.Generator (masks from torch's
    CPU stream, like the reference's nn.Dropout but not reproducible elsewhere) or an int stream key (`call_key(seed, call)`):
    then site k of the call uses counter_keep_mask(numel, keep, key, k) over the call's own tensor (row-major element index),
    which a native kernel regenerates bit for bit."""

    def __init__(self, cfg: BertCfg, generator, enabled: bool):
        self.cfg, self.gen, self.enabled = cfg, generator, enabled
        self.site = 0

    def __call__(self, x: Tensor, p: float) -> Tensor:
        site = self.site
        self.site += 1            # site numbers are positions in the forward, whether or not that dropout is active
        if not self.enabled or p == 0.0:
            return x
        keep = 1.0 - p
        if isinstance(self.gen, int):
            m = counter_keep_mask(x.numel(), keep, self.gen, site).view(x.shape).to(x.dtype).div_(keep)
        else:
            m = torch.empty_like(x).bernoulli_(keep, generator=self.gen).div_(keep)
        return x * m


def bert_forward(p: Dict[str, Tensor], x: Dict[str, Tensor], cfg: BertCfg, drop: Optional[BertDropout] = None):
    """-> (logits [B, C], feat [B, 768]).  x = {'input_ids': int64 [B, L], 'attention_mask': int64 [B, L]} (bert.py:34)."""
    ids, am = x["input_ids"], x.get("attention_mask")
    B, L = ids.shape
    H, nh = cfg.hidden, cfg.heads
    dh = H // nh
    drop = drop or BertDropout(cfg, None, False)
    # nn.Embedding(vocab, hidden, padding_idx=pad_token_id = 0): row 0 (the padding token) never receives a gradient
    e = F.embedding(ids, p["bert.embeddings.word_embeddings.weight"], padding_idx=0) + p["bert.embeddings.token_type_embeddings.weight"][0]
    e = e + p["bert.embeddings.position_embeddings.weight"][:L]
    h = F.layer_norm(e, (H,), p["bert.embeddings.LayerNorm.weight"], p["bert.embeddings.LayerNorm.bias"], cfg.eps)
    h = drop(h, cfg.hidden_dropout)
    keymask = None
    if am is not None and not bool(am.all()):
        # additive key-padding mask: finfo.min on padded keys (exp underflows to exactly 0), 0 elsewhere
        keymask = torch.zeros(B, 1, 1, L, dtype=h.dtype).masked_fill(am[:, None, None, :] == 0, torch.finfo(h.dtype).min)
    for i in range(cfg.layers):
        pre = f"bert.encoder.layer.{i}."
        q = F.linear(h, p[pre + "attention.self.query.weight"], p[pre + "attention.self.query.bias"]).view(B, L, nh, dh).transpose(1, 2)
        k = F.linear(h, p[pre + "attention.self.key.weight"], p[pre + "attention.self.key.bias"]).view(B, L, nh, dh).transpose(1, 2)
        v = F.linear(h, p[pre + "attention.self.value.weight"], p[pre + "attention.self.value.bias"]).view(B, L, nh, dh).transpose(1, 2)
        s = torch.matmul(q, k.transpose(2, 3)) * (dh ** -0.5)
        if keymask is not None:
            s = s + keymask
        a = drop(F.softmax(s, dim=-1), cfg.attn_dropout)
        ctx = torch.matmul(a, v).transpose(1, 2).contiguous().reshape(B, L, H)
        o = drop(F.linear(ctx, p[pre + "attention.output.dense.weight"], p[pre + "attention.output.dense.bias"]), cfg.hidden_dropout)
        h = F.layer_norm(o + h, (H,), p[pre + "attention.output.LayerNorm.weight"], p[pre + "attention.output.LayerNorm.bias"], cfg.eps)
        f = F.gelu(F.linear(h, p[pre + "intermediate.dense.weight"], p[pre + "intermediate.dense.bias"]))
        f = drop(F.linear(f, p[pre + "output.dense.weight"], p[pre + "output.dense.bias"]), cfg.hidden_dropout)
        h = F.layer_norm(f + h, (H,), p[pre + "output.LayerNorm.weight"], p[pre + "output.LayerNorm.bias"], cfg.eps)
    feat = torch.mean(drop(h, cfg.pooled_dropout), 1)                       # over all L positions, padding included
    z = F.gelu(F.linear(feat, p["classifier.0.weight"], p["classifier.0.bias"]))
    return F.linear(z, p["classifier.2.weight"], p["classifier.2.bias"]), feat


def bert_layer_id(name: str, layers: int) -> int:
    """group_matcher of bert.py:58-60 through group_with_matcher (nets/utils.py:207-268): embeddings -> 0,
    encoder.layer.i -> i + 1, everything unmatched (pooler, classifier) -> layers + 1."""
    if name.startswith("bert.embeddings"):
        return 0
    if name.startswith("bert.encoder.layer."):
        return int(name.split(".")[3]) + 1
    return layers + 1


def bert_param_hparams(names_shapes, layers: int, lr: float, weight_decay: float, layer_decay: float):
    """{name: (lr, weight_decay)} as param_groups_layer_decay builds them (nets/utils.py:143-204): 1-D tensors are not
    decayed, `no_weight_decay()` is empty for BERT (bert.py:62-63), lr scale = layer_decay ** (layers + 1 - layer id)."""
    out = {}
    for n, shp in names_shapes:
        scale = layer_decay ** (layers + 1 - bert_layer_id(n, layers)) if layer_decay != 1.0 else 1.0
        out[n] = (scale * lr, 0.0 if len(shp) == 1 else weight_decay)
    return out


class BertSSLOracle(O.SSLOracle):
    """The SSL step of ssl_oracle.SSLOracle with the text backbone and `use_cat: False` (srsoftmatch.py:119-130,
    srflexmatch.py:119-130, ...): labelled and strong batches go through the model with autograd, the weak batch under
    no_grad — three calls, so dropout (when on) draws masks in that order."""

    def __init__(self, bert_cfg: BertCfg, cfg: O.StepConfig, params, rewarder, generator, stochastic: bool = False):
        hp = bert_param_hparams(bert_cfg.param_shapes(), bert_cfg.layers, cfg.lr, cfg.weight_decay, cfg.layer_decay)
        super().__init__(None, cfg, params, rewarder, generator, hparams=hp)
        self.bert_cfg = bert_cfg
        self.stochastic = stochastic
        self.calls = 0          # backbone calls made so far (counter-dropout stream index)

    def _drop(self):
        """Dropout source of the next backbone call: `drop_gen` a torch.Generator -> torch's stream; an int seed -> the counter
        masks with one stream key per call (`call_key(seed, number of calls so far)`)."""
        if isinstance(self.drop_gen, int):
            d = BertDropout(self.bert_cfg, call_key(self.drop_gen, self.calls), self.stochastic)
        else:
            d = BertDropout(self.bert_cfg, self.drop_gen, self.stochastic)
        self.calls += 1
        return d

    def _backbone(self, x_lb, x_ulb_w, x_ulb_s):
        # srsoftmatch.py:119-130 order: labelled, strong (both with autograd), then weak under no_grad
        llb, flb = bert_forward(self.p, x_lb, self.bert_cfg, self._drop())
        ls, fs = bert_forward(self.p, x_ulb_s, self.bert_cfg, self._drop())
        with torch.no_grad():
            lw, fw = bert_forward(self.p, x_ulb_w, self.bert_cfg, self._drop())
        return llb, lw, ls, flb, fw, fs


def build_det_bert_oracle(bert_cfg: BertCfg, cfg: O.StepConfig, seed: int = 0, head_gain: float = 1.0, stochastic: bool = False) -> BertSSLOracle:
    """Initialised from semireward_b200.detgen fills, like ssl_oracle.build_det_oracle (same tensors the golden generator
    loads into the live reference)."""
    from semireward_b200 import detgen
    p = {n: torch.from_numpy(detgen.fill_param(n, s, seed)) for n, s in bert_cfg.param_shapes()}
    if head_gain != 1.0:
        p["classifier.2.weight"] = p["classifier.2.weight"] * head_gain
    rp = {n: torch.from_numpy(detgen.fill_param("rewarder." + n, s, seed)) for n, s in O.rewarder_param_shapes(cfg.feature_dim, cfg.num_classes)}
    gp = {n: torch.from_numpy(detgen.fill_param("generator." + n, s, seed)) for n, s in O.generator_param_shapes(cfg.feature_dim)}
    return BertSSLOracle(bert_cfg, cfg, p, rp, gp, stochastic=stochastic)


def check_flops():
    assert abs(BertCfg().fwd_flops_per_seq(512) / 1e9 - 96.64) < 0.05, BertCfg().fwd_flops_per_seq(512)
    return math.isfinite(1.0)
