"""TEST INFRASTRUCTURE — CPU restatement of the audio path of the hot loop (BASELINE configs[4]):
`ClassificationHubert` (semilearn/nets/hubert/hubert.py:10-50) around a HuBERT encoder, driven by the SSL step of
oracle/ssl_oracle.py with `use_cat: False`.

Only tests/ (and, later, smoke()/bench.py's CPU arm) may import this module; the product never does.

Third-party arithmetic: the encoder is Hugging Face `transformers.HubertModel` (`transformers>=4.30.0` in
the reference's requirements.txt, unpinned; 5.5.0 installed here = the de-facto pin), not under /root/reference.  Its
published algorithm (Hsu et al. 2021; modeling_hubert.py of the installed version: `feat_extract_norm='group'`,
`do_stable_layer_norm=False`, eager attention) is restated here in plain torch ops in the same order and pinned against
the live `ClassificationHubert` on CPU (tests/test_hubert_oracle.py, fixtures tests/golden/hubert_*.npz made by
tests/golden/make_golden_hubert.py).

One forward (B clips of T samples, 16 kHz):
  conv stem: 7 Conv1d without bias, 512 channels, kernels (10,3,3,3,3,2,2), strides (5,2,2,2,2,2,2), GELU after each;
             GroupNorm(512 groups) after the first conv only                         -> [B, 512, F]   (F = 199 for 64 000 samples)
  projection: LayerNorm(512) -> Linear(512, 768) -> dropout(feat_proj_dropout = 0)
  [SpecAugment: in train mode 5 % of the frames (spans of 10, >= 2 spans) are replaced by `masked_spec_embed`]
  positional conv: Conv1d(768, 768, k = 128, pad 64, groups 16) with weight norm over (out, in) per tap, last frame dropped,
             GELU; h = LayerNorm(h + pos) -> dropout
  12 post-LN layers, each skipped with probability `layerdrop` = 0.1 in train mode:
             h = LN(h + drop(attn(h)));  h = LN(h + drop(W2 drop(gelu(W1 h))))
  feat = mean over the F frames of drop(h);  logits = gelu(feat Wc1 + bc1) Wc2 + bc2        (hubert.py:17-21, 44-49)
The reference passes NO attention mask (hubert.py:45), so padded clips attend everywhere.

Deterministic parity mode = every dropout 0, layerdrop 0, apply_spec_augment False: the four sources of randomness of a
train-mode pass (dropout x5 sites, LayerDrop, SpecAugment spans from numpy's global RNG, and the wrapper's own dropout)
cannot be reproduced outside the reference's RNG streams; a stochastic pass can only be compared statistically.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import ssl_oracle as O

Tensor = torch.Tensor


@dataclass
class HubertCfg:
    hidden: int = 768                   # ClassificationHubert hard-codes 768 (hubert.py:17-21)
    layers: int = 12
    heads: int = 12
    intermediate: int = 3072
    conv_dim: Tuple[int, ...] = (512,) * 7
    conv_kernel: Tuple[int, ...] = (10, 3, 3, 3, 3, 2, 2)
    conv_stride: Tuple[int, ...] = (5, 2, 2, 2, 2, 2, 2)
    pos_kernel: int = 128
    pos_groups: int = 16
    eps: float = 1e-5
    num_classes: int = 10
    # train-mode dropout probabilities (facebook/hubert-base-ls960: 0.1 each, layerdrop 0.1); 0 = deterministic parity mode
    feat_proj_dropout: float = 0.0
    hidden_dropout: float = 0.0
    attention_dropout: float = 0.0
    activation_dropout: float = 0.0
    pooled_dropout: float = 0.0          # hubert.py:15 (the wrapper's own nn.Dropout(0.1))

    def frames(self, samples: int) -> int:
        n = samples
        for k, s in zip(self.conv_kernel, self.conv_stride):
            n = (n - k) // s + 1
        return n

    def param_shapes(self) -> List[Tuple[str, Tuple[int, ...]]]:
        """`ClassificationHubert.named_parameters()` order."""
        H, I, C = self.hidden, self.intermediate, self.conv_dim
        out = [("model.masked_spec_embed", (H,)), ("model.feature_extractor.conv_layers.0.conv.weight", (C[0], 1, self.conv_kernel[0])),
               ("model.feature_extractor.conv_layers.0.layer_norm.weight", (C[0],)), ("model.feature_extractor.conv_layers.0.layer_norm.bias", (C[0],))]
        for i in range(1, len(C)):
            out.append((f"model.feature_extractor.conv_layers.{i}.conv.weight", (C[i], C[i - 1], self.conv_kernel[i])))
        out += [("model.feature_projection.layer_norm.weight", (C[-1],)), ("model.feature_projection.layer_norm.bias", (C[-1],)),
                ("model.feature_projection.projection.weight", (H, C[-1])), ("model.feature_projection.projection.bias", (H,)),
                ("model.encoder.pos_conv_embed.conv.bias", (H,)),
                ("model.encoder.pos_conv_embed.conv.parametrizations.weight.original0", (1, 1, self.pos_kernel)),
                ("model.encoder.pos_conv_embed.conv.parametrizations.weight.original1", (H, H // self.pos_groups, self.pos_kernel)),
                ("model.encoder.layer_norm.weight", (H,)), ("model.encoder.layer_norm.bias", (H,))]
        for i in range(self.layers):
            p = f"model.encoder.layers.{i}."
            for proj in ("k_proj", "v_proj", "q_proj", "out_proj"):
                out += [(p + f"attention.{proj}.weight", (H, H)), (p + f"attention.{proj}.bias", (H,))]
            out += [(p + "layer_norm.weight", (H,)), (p + "layer_norm.bias", (H,)),
                    (p + "feed_forward.intermediate_dense.weight", (I, H)), (p + "feed_forward.intermediate_dense.bias", (I,)),
                    (p + "feed_forward.output_dense.weight", (H, I)), (p + "feed_forward.output_dense.bias", (H,)),
                    (p + "final_layer_norm.weight", (H,)), (p + "final_layer_norm.bias", (H,))]
        out += [("classifier.0.weight", (H, H)), ("classifier.0.bias", (H,)), ("classifier.2.weight", (self.num_classes, H)),
                ("classifier.2.bias", (self.num_classes,))]
        return out

    def fwd_flops_per_clip(self, samples: int = 64000) -> float:
        """2*MACs of one clip: conv stem + projection + positional conv + encoder + classifier (~56.9 GF for 4 s)."""
        n, fl, cin = samples, 0.0, 1
        for c, k, s in zip(self.conv_dim, self.conv_kernel, self.conv_stride):
            n = (n - k) // s + 1
            fl += 2.0 * n * c * cin * k
            cin = c
        H, I = self.hidden, self.intermediate
        fl += 2.0 * n * cin * H
        fl += 2.0 * n * H * (H // self.pos_groups) * self.pos_kernel
        fl += self.layers * (2.0 * n * (4 * H * H + 2 * H * I) + 4.0 * n * n * H)
        return fl + 2.0 * (H * H + H * self.num_classes)


class HubertDropout:
    """Counter-based masks of one model call (include/srw.h `srw_dropout`; bert_oracle.counter_keep_mask): site 0 = feature
    projection output [B, F, H], 1 = encoder input, 2 + 4 l = attention probabilities of layer l [B, heads, F, F], 3 + 4 l =
    attention output, 4 + 4 l = FFN activation [B, F, I], 5 + 4 l = FFN output, 2 + 4 layers = the wrapper's dropout before the
    mean pool.  A layer left out by LayerDrop keeps its site numbers.  `key` None = no dropout."""

    def __init__(self, key):
        self.key = key

    def __call__(self, x: Tensor, p: float, site: int) -> Tensor:
        if self.key is None or p == 0.0:
            return x
        from .bert_oracle import counter_keep_mask
        keep = 1.0 - p
        return x * counter_keep_mask(x.numel(), keep, self.key, site).view(x.shape).to(x.dtype).div_(keep)


def hubert_forward(p: Dict[str, Tensor], x: Tensor, cfg: HubertCfg, mask_time_indices: Optional[Tensor] = None, skip_layers=(), drop_key=None):
    """-> (logits [B, C], feat [B, 768]).  x: fp32 [B, T].  Default = deterministic parity mode (no dropout / LayerDrop /
    SpecAugment).  The sources of randomness of a train-mode pass are explicit inputs: `mask_time_indices` bool [B, F]
    (SpecAugment: those frames are replaced by `masked_spec_embed` after the feature projection, modeling_hubert.py
    `_mask_hidden_states`), `skip_layers` (LayerDrop: encoder layers left out) and `drop_key` (an int stream key: every
    nn.Dropout of the call becomes the counter mask of HubertDropout with the cfg's probabilities)."""
    H, nh = cfg.hidden, cfg.heads
    dh = H // nh
    drop = HubertDropout(drop_key)
    h = x[:, None]
    for i in range(len(cfg.conv_dim)):
        pre = f"model.feature_extractor.conv_layers.{i}."
        h = F.conv1d(h, p[pre + "conv.weight"], None, stride=cfg.conv_stride[i])
        if i == 0:
            h = F.group_norm(h, cfg.conv_dim[0], p[pre + "layer_norm.weight"], p[pre + "layer_norm.bias"], 1e-5)
        h = F.gelu(h)
    h = h.transpose(1, 2)                                                     # [B, F, 512]
    h = F.layer_norm(h, (cfg.conv_dim[-1],), p["model.feature_projection.layer_norm.weight"], p["model.feature_projection.layer_norm.bias"], cfg.eps)
    h = F.linear(h, p["model.feature_projection.projection.weight"], p["model.feature_projection.projection.bias"])
    h = drop(h, cfg.feat_proj_dropout, 0)
    B, Fr, _ = h.shape
    if mask_time_indices is not None:
        h = torch.where(mask_time_indices[..., None], p["model.masked_spec_embed"].to(h.dtype), h)
    # positional convolution, weight-normalised over (out, in) for every tap: w = g * v / ||v||  (weight_norm(dim=2))
    g = p["model.encoder.pos_conv_embed.conv.parametrizations.weight.original0"]
    v = p["model.encoder.pos_conv_embed.conv.parametrizations.weight.original1"]
    w = torch._weight_norm(v, g, 2)
    pos = F.conv1d(h.transpose(1, 2), w, p["model.encoder.pos_conv_embed.conv.bias"], padding=cfg.pos_kernel // 2, groups=cfg.pos_groups)
    if cfg.pos_kernel % 2 == 0:
        pos = pos[:, :, :-1]
    pos = F.gelu(pos).transpose(1, 2)
    h = F.layer_norm(h + pos, (H,), p["model.encoder.layer_norm.weight"], p["model.encoder.layer_norm.bias"], cfg.eps)
    h = drop(h, cfg.hidden_dropout, 1)
    for i in range(cfg.layers):
        if i in skip_layers:
            continue
        pre = f"model.encoder.layers.{i}."
        q = F.linear(h, p[pre + "attention.q_proj.weight"], p[pre + "attention.q_proj.bias"]).view(B, Fr, nh, dh).transpose(1, 2)
        k = F.linear(h, p[pre + "attention.k_proj.weight"], p[pre + "attention.k_proj.bias"]).view(B, Fr, nh, dh).transpose(1, 2)
        vv = F.linear(h, p[pre + "attention.v_proj.weight"], p[pre + "attention.v_proj.bias"]).view(B, Fr, nh, dh).transpose(1, 2)
        a = drop(F.softmax(torch.matmul(q, k.transpose(2, 3)) * (dh ** -0.5), dim=-1), cfg.attention_dropout, 2 + 4 * i)
        ctx = torch.matmul(a, vv).transpose(1, 2).contiguous().reshape(B, Fr, H)
        o = drop(F.linear(ctx, p[pre + "attention.out_proj.weight"], p[pre + "attention.out_proj.bias"]), cfg.hidden_dropout, 3 + 4 * i)
        h = F.layer_norm(h + o, (H,), p[pre + "layer_norm.weight"], p[pre + "layer_norm.bias"], cfg.eps)
        f = F.gelu(F.linear(h, p[pre + "feed_forward.intermediate_dense.weight"], p[pre + "feed_forward.intermediate_dense.bias"]))
        f = drop(f, cfg.activation_dropout, 4 + 4 * i)
        f = drop(F.linear(f, p[pre + "feed_forward.output_dense.weight"], p[pre + "feed_forward.output_dense.bias"]), cfg.hidden_dropout, 5 + 4 * i)
        h = F.layer_norm(h + f, (H,), p[pre + "final_layer_norm.weight"], p[pre + "final_layer_norm.bias"], cfg.eps)
    feat = torch.mean(drop(h, cfg.pooled_dropout, 2 + 4 * cfg.layers), 1)
    z = F.gelu(F.linear(feat, p["classifier.0.weight"], p["classifier.0.bias"]))
    return F.linear(z, p["classifier.2.weight"], p["classifier.2.bias"]), feat


def hubert_layer_id(name: str, layers: int) -> int:
    """group_matcher of hubert.py:51-53 through group_with_matcher (nets/utils.py:207-268): conv stem, feature projection and
    positional conv -> 0, encoder.layers.i -> i + 1, everything unmatched (masked_spec_embed, encoder.layer_norm, classifier)
    -> layers + 1."""
    if name.startswith(("model.feature_projection", "model.feature_extractor", "model.encoder.pos_conv_embed")):
        return 0
    if name.startswith("model.encoder.layers."):
        return int(name.split(".")[3]) + 1
    return layers + 1


def hubert_param_hparams(names_shapes, layers: int, lr: float, weight_decay: float, layer_decay: float):
    """{name: (lr, weight_decay)} as param_groups_layer_decay builds them: 1-D tensors are not decayed (masked_spec_embed and
    all biases / norms; the weight-norm gain is 3-D [1,1,128] and IS decayed), no_weight_decay() is empty (hubert.py:55-56)."""
    out = {}
    for n, shp in names_shapes:
        scale = layer_decay ** (layers + 1 - hubert_layer_id(n, layers)) if layer_decay != 1.0 else 1.0
        out[n] = (scale * lr, 0.0 if len(shp) == 1 else weight_decay)
    return out


class HubertSSLOracle(O.SSLOracle):
    """ssl_oracle.SSLOracle with the audio backbone and `use_cat: False` (three calls; the weak one under no_grad).
    `drop_seed` (int) switches the counter-based dropout on: call c of the run uses stream key call_key(drop_seed, c).
    `stochastic_inputs(call) -> (mask_time_indices, skip_layers)` injects SpecAugment / LayerDrop decisions per call."""

    def __init__(self, hubert_cfg: HubertCfg, cfg: O.StepConfig, params, rewarder, generator):
        hp = hubert_param_hparams(hubert_cfg.param_shapes(), hubert_cfg.layers, cfg.lr, cfg.weight_decay, cfg.layer_decay)
        super().__init__(None, cfg, params, rewarder, generator, hparams=hp)
        self.hubert_cfg = hubert_cfg
        self.drop_seed = None
        self.calls = 0
        self.stochastic_inputs = None

    def _call(self, x):
        from .bert_oracle import call_key
        key = None if self.drop_seed is None else call_key(self.drop_seed, self.calls)
        mti, skip = (None, ()) if self.stochastic_inputs is None else self.stochastic_inputs(self.calls)
        self.calls += 1
        return hubert_forward(self.p, x, self.hubert_cfg, mask_time_indices=mti, skip_layers=skip, drop_key=key)

    def _backbone(self, x_lb, x_ulb_w, x_ulb_s):
        llb, flb = self._call(x_lb)
        ls, fs = self._call(x_ulb_s)
        with torch.no_grad():
            lw, fw = self._call(x_ulb_w)
        return llb, lw, ls, flb, fw, fs


def build_det_hubert_oracle(hubert_cfg: HubertCfg, cfg: O.StepConfig, seed: int = 0, head_gain: float = 1.0) -> HubertSSLOracle:
    from semireward_b200 import detgen
    p = {n: torch.from_numpy(detgen.fill_param(n, s, seed)) for n, s in hubert_cfg.param_shapes()}
    if head_gain != 1.0:
        p["classifier.2.weight"] = p["classifier.2.weight"] * head_gain
    rp = {n: torch.from_numpy(detgen.fill_param("rewarder." + n, s, seed)) for n, s in O.rewarder_param_shapes(cfg.feature_dim, cfg.num_classes)}
    gp = {n: torch.from_numpy(detgen.fill_param("generator." + n, s, seed)) for n, s in O.generator_param_shapes(cfg.feature_dim)}
    return HubertSSLOracle(hubert_cfg, cfg, p, rp, gp)
