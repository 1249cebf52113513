"""TEST INFRASTRUCTURE — CPU restatement (the oracle) of SemiReward's per-step SSL hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product path (semireward_b200/) never does and fails loudly when its CUDA library is missing.

What is restated (fp32, torch CPU, functional style — no nn.Module of the reference is copied):
  * ViT backbone forward            — /root/reference/semilearn/nets/vit/vit.py:39-44, 69-75, 91-107, 163-166, 277-306
  * timm DropPath (un-vendored dep) — SURVEY.md §8c: per-sample Bernoulli(keep)/keep, identity in eval
  * Rewarder / Generator forward    — semilearn/algorithms/semireward/semireward.py:52-72, 21-24
  * cosine_similarity_n, label_dim  — semireward.py:130-139, 147-148
  * FlexMatch / FreeMatch / SoftMatch masking hooks, DistAlign EMA
                                    — algorithms/srflexmatch/utils.py:23-63, freematch/utils.py:23-66,
                                      srsoftmatch/utils.py:31-77, hooks/dist_align.py:25-55
  * pseudo-labels                   — algorithms/hooks/pseudo_label.py:16-52
  * ce / consistency loss           — core/criterions/cross_entropy.py:11-31, consistency.py:13-45
  * FreeMatch fairness entropy loss — algorithms/srfreematch/srfreematch.py:16-44
  * train_step / data_generator     — srflexmatch.py:72-217, srfreematch.py:76-228, srsoftmatch.py:61-221,
                                      srfixmatch/fixmatch.py:62-210 (= srflexmatch with the stateless FixedThresholdingHook),
                                      srpseudolabel/srpseudolabel.py:58-201
  * sr_decay                        — core/algorithmbase.py:177-183
  * ParamUpdateHook                 — core/hooks/param_update.py:21-40 (backward, AdamW, LambdaLR, zero_grad)
  * AdamW groups + cosine schedule  — core/utils/build.py:193-251, nets/utils.py:143-204 (layer decay)
  * Adam for the Rewarder           — srflexmatch.py:54 (torch.optim.Adam defaults)

Parity pin: the reference ships NO tests or golden vectors, so this oracle is pinned against
outputs of the live reference itself, run in the build container by tests/golden/make_golden.py (fixtures in
tests/golden/*.npz) and re-checked live by tests/test_oracle_golden.py (test_oracle_matches_live_reference) whenever /root/reference is mounted.
Gradients come from torch autograd over this functional forward (autograd is the CPU differentiation engine of
the reference as well).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------------
# Backbone
# ------------------------------------------------------------------------------------------------
@dataclass
class ViTConfig:
    img_size: int = 32
    patch_size: int = 2
    in_chans: int = 3
    embed_dim: int = 384
    depth: int = 12
    num_heads: int = 6
    mlp_ratio: float = 4.0
    num_classes: int = 100
    drop_path_rate: float = 0.0
    ln_eps: float = 1e-6  # vit.py:222

    @property
    def num_patches(self) -> int:
        return (self.img_size // self.patch_size) ** 2

    @property
    def tokens(self) -> int:
        return self.num_patches + 1

    @property
    def hidden(self) -> int:
        return int(self.embed_dim * self.mlp_ratio)

    def param_shapes(self) -> List[Tuple[str, Tuple[int, ...]]]:
        """state_dict order and shapes of the reference VisionTransformer (vit.py:232-275; 152 tensors at depth 12)."""
        D, P, C = self.embed_dim, self.patch_size, self.in_chans
        out = [("cls_token", (1, 1, D)), ("pos_embed", (1, self.tokens, D)),
               ("patch_embed.proj.weight", (D, C, P, P)), ("patch_embed.proj.bias", (D,))]
        for i in range(self.depth):
            b = f"blocks.{i}."
            out += [(b + "norm1.weight", (D,)), (b + "norm1.bias", (D,)),
                    (b + "attn.qkv.weight", (3 * D, D)), (b + "attn.qkv.bias", (3 * D,)),
                    (b + "attn.proj.weight", (D, D)), (b + "attn.proj.bias", (D,)),
                    (b + "norm2.weight", (D,)), (b + "norm2.bias", (D,)),
                    (b + "mlp.fc1.weight", (self.hidden, D)), (b + "mlp.fc1.bias", (self.hidden,)),
                    (b + "mlp.fc2.weight", (D, self.hidden)), (b + "mlp.fc2.bias", (D,))]
        out += [("norm.weight", (D,)), ("norm.bias", (D,)),
                ("head.weight", (self.num_classes, D)), ("head.bias", (self.num_classes,))]
        return out

    def drop_path_rates(self) -> List[float]:
        # vit.py:247-249  dpr = linspace(0, drop_path_rate, depth)
        return [float(v) for v in torch.linspace(0, self.drop_path_rate, self.depth)]


def draw_drop_path_masks(cfg: ViTConfig, batch: int, generator: Optional[torch.Generator] = None) -> Optional[Tensor]:
    """[depth, 2, batch] multipliers mask/keep (timm DropPath).  None when the rate is 0."""
    if cfg.drop_path_rate == 0.0:
        return None
    rates = cfg.drop_path_rates()
    m = torch.ones(cfg.depth, 2, batch)
    for i, r in enumerate(rates):
        if r > 0.0:
            keep = 1.0 - r
            m[i] = torch.empty(2, batch).bernoulli_(keep, generator=generator) / keep
    return m


def vit_forward(p: Dict[str, Tensor], x: Tensor, cfg: ViTConfig, drop_masks: Optional[Tensor] = None, return_tokens: bool = False):
    """logits [B,C], feat [B,D] = CLS token after the final LayerNorm (vit.py:277-306, global_pool='token').
    return_tokens: also the whole normalised token matrix [B,N,D] = VisionTransformer.extract (vit.py:277-283)."""
    B = x.shape[0]
    D, H = cfg.embed_dim, cfg.num_heads
    dh = D // H
    scale = dh ** -0.5
    # PatchEmbed: conv k=s=P, flatten(2).transpose(1,2)  (vit.py:39-44)
    t = F.conv2d(x, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"], stride=cfg.patch_size)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((p["cls_token"].expand(B, -1, -1), t), dim=1) + p["pos_embed"]  # vit.py:279-280
    N = t.shape[1]
    for i in range(cfg.depth):
        b = f"blocks.{i}."
        # attention branch (vit.py:91-107, 164)
        y = F.layer_norm(t, (D,), p[b + "norm1.weight"], p[b + "norm1.bias"], cfg.ln_eps)
        qkv = F.linear(y, p[b + "attn.qkv.weight"], p[b + "attn.qkv.bias"]).reshape(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
        o = (attn @ v).transpose(1, 2).reshape(B, N, D)
        o = F.linear(o, p[b + "attn.proj.weight"], p[b + "attn.proj.bias"])
        if drop_masks is not None:
            o = o * drop_masks[i, 0].view(B, 1, 1)
        t = t + o
        # MLP branch (vit.py:69-75, 165); nn.GELU default = exact erf
        y = F.layer_norm(t, (D,), p[b + "norm2.weight"], p[b + "norm2.bias"], cfg.ln_eps)
        h = F.gelu(F.linear(y, p[b + "mlp.fc1.weight"], p[b + "mlp.fc1.bias"]))
        h = F.linear(h, p[b + "mlp.fc2.weight"], p[b + "mlp.fc2.bias"])
        if drop_masks is not None:
            h = h * drop_masks[i, 1].view(B, 1, 1)
        t = t + h
    t = F.layer_norm(t, (D,), p["norm.weight"], p["norm.bias"], cfg.ln_eps)
    feat = t[:, 0]
    logits = F.linear(feat, p["head.weight"], p["head.bias"])
    if return_tokens:
        return logits, feat, t
    return logits, feat


# ------------------------------------------------------------------------------------------------
# SemiReward modules
# ------------------------------------------------------------------------------------------------
def label_dim(num_classes: int, default_dim: int = 100) -> int:
    return int(max(default_dim, num_classes))  # semireward.py:147-148


def rewarder_param_shapes(feature_dim: int, num_classes: int, emb: int = 128):
    L = label_dim(num_classes)
    return [("feature_fc.weight", (128, feature_dim)), ("feature_fc.bias", (128,)),
            ("feature_norm.weight", (128,)), ("feature_norm.bias", (128,)),
            ("label_embedding.weight", (L, emb)), ("label_norm.weight", (emb,)), ("label_norm.bias", (emb,)),
            ("cross_attention_fc.weight", (1, 128)), ("cross_attention_fc.bias", (1,)),
            ("mlp_fc1.weight", (256, 128)), ("mlp_fc1.bias", (256,)),
            ("mlp_fc2.weight", (128, 256)), ("mlp_fc2.bias", (128,)),
            ("ffn_fc1.weight", (64, 128)), ("ffn_fc1.bias", (64,)),
            ("ffn_fc2.weight", (1, 64)), ("ffn_fc2.bias", (1,))]


def generator_param_shapes(feature_dim: int):
    return [("fc_layers.0.weight", (256, feature_dim)), ("fc_layers.0.bias", (256,)),
            ("fc_layers.2.weight", (128, 256)), ("fc_layers.2.bias", (128,)),
            ("fc_layers.4.weight", (64, 128)), ("fc_layers.4.bias", (64,)),
            ("fc_layers.6.weight", (1, 64)), ("fc_layers.6.bias", (1,))]


def rewarder_forward(rp: Dict[str, Tensor], feats: Tensor, labels: Tensor) -> Tensor:
    """reward [B,1] in (0,1).  semireward.py:52-72.  NB the softmax runs over the 2B rows (dim=0)."""
    f = F.layer_norm(F.linear(feats, rp["feature_fc.weight"], rp["feature_fc.bias"]), (128,),
                     rp["feature_norm.weight"], rp["feature_norm.bias"], 1e-5)
    e = F.layer_norm(rp["label_embedding.weight"][labels], (rp["label_embedding.weight"].shape[1],),
                     rp["label_norm.weight"], rp["label_norm.bias"], 1e-5)
    X = torch.cat((f, e), dim=0)
    w = torch.softmax(F.linear(X, rp["cross_attention_fc.weight"], rp["cross_attention_fc.bias"]), dim=0)
    ctx = (w * X).sum(dim=0)
    h = ctx.unsqueeze(0).expand(e.size(0), -1) + e
    h = F.linear(F.relu(F.linear(h, rp["mlp_fc1.weight"], rp["mlp_fc1.bias"])), rp["mlp_fc2.weight"], rp["mlp_fc2.bias"])
    h = F.relu(F.linear(h, rp["ffn_fc1.weight"], rp["ffn_fc1.bias"]))
    return torch.sigmoid(F.linear(h, rp["ffn_fc2.weight"], rp["ffn_fc2.bias"]))


def generator_forward(gp: Dict[str, Tensor], feats: Tensor) -> Tensor:
    """[B,1] >= 0.  semireward.py:9-24 (ReLU after the last Linear as well)."""
    h = feats
    for i in (0, 2, 4):
        h = F.relu(F.linear(h, gp[f"fc_layers.{i}.weight"], gp[f"fc_layers.{i}.bias"]))
    return F.relu(F.linear(h, gp["fc_layers.6.weight"], gp["fc_layers.6.bias"]))


def sr_target(gen_label: Tensor, true_label: Tensor, num_classes: int) -> Tensor:
    """cosine_similarity_n(one_hot(gen), one_hot(true)) -> [B,1] in {1.0, 0.5} (semireward.py:130-139,
    srflexmatch.py:195-197).  Restated literally (one-hot + cosine) so that the analytic shortcut used by the CUDA
    path (equal ? 1 : 0.5) is itself checked against it."""
    a = F.one_hot(gen_label, num_classes=num_classes).float()
    b = F.one_hot(true_label, num_classes=num_classes).float()
    cos = torch.cosine_similarity(a, b, dim=-1, eps=1e-8)
    return ((cos + 1) / 2).view(a.size(0), 1)


# ------------------------------------------------------------------------------------------------
# Losses
# ------------------------------------------------------------------------------------------------
def ce_loss(logits: Tensor, targets: Tensor, reduction: str = "none") -> Tensor:
    logp = F.log_softmax(logits, dim=-1)  # cross_entropy.py:29-31 (hard-label branch)
    return F.nll_loss(logp, targets, reduction=reduction)


def consistency_loss(logits: Tensor, targets: Tensor, mask: Optional[Tensor] = None, mask2: Optional[Tensor] = None) -> Tensor:
    loss = ce_loss(logits, targets, "none")  # consistency.py:38 ('ce')
    if mask is not None:
        loss = loss * mask
    if mask2 is not None:
        loss = loss * mask2
    return loss.mean()  # mean over B, not over sum(mask)  (consistency.py:45)


def freematch_entropy_loss(mask: Tensor, logits_s: Tensor, prob_model: Tensor, label_hist: Tensor):
    """srfreematch.py:16-44."""
    sel = logits_s[mask.bool()]
    prob_s = sel.softmax(dim=-1)
    pred = prob_s.argmax(dim=-1)
    hist_s = torch.bincount(pred, minlength=sel.shape[1]).to(sel.dtype)
    hist_s = hist_s / hist_s.sum()
    inv = 1 / label_hist.reshape(1, -1)
    inv = torch.where(torch.isinf(inv) & (inv > 0), torch.zeros_like(inv), inv).detach()
    mod_prob_model = prob_model.reshape(1, -1) * inv
    mod_prob_model = mod_prob_model / mod_prob_model.sum(dim=-1, keepdim=True)
    inv_s = 1 / hist_s
    inv_s = torch.where(torch.isinf(inv_s) & (inv_s > 0), torch.zeros_like(inv_s), inv_s).detach()
    mod_mean = prob_s.mean(dim=0, keepdim=True) * inv_s
    mod_mean = mod_mean / mod_mean.sum(dim=-1, keepdim=True)
    loss = (mod_prob_model * torch.log(mod_mean + 1e-12)).sum(dim=1)
    return loss.mean()


# ------------------------------------------------------------------------------------------------
# Hook state machines
# ------------------------------------------------------------------------------------------------
class FlexMatchState:
    """FlexMatchThresholdingHook (srflexmatch/utils.py:11-63).  The reference rebuilds a host Counter over all
    ulb_dest_len entries at every call; the restatement keeps the identical integer histogram with bincount."""

    def __init__(self, ulb_dest_len: int, num_classes: int, thresh_warmup: bool = True):
        self.ulb_dest_len, self.num_classes, self.thresh_warmup = ulb_dest_len, num_classes, thresh_warmup
        self.selected_label = torch.full((ulb_dest_len,), -1, dtype=torch.long)
        self.classwise_acc = torch.zeros(num_classes)

    def update(self):
        counts = torch.bincount(self.selected_label + 1, minlength=self.num_classes + 1)  # bucket 0 == label -1
        max_all = int(counts.max())
        if max_all < self.ulb_dest_len:  # utils.py:26
            if self.thresh_warmup:
                denom = max_all  # the -1 bucket takes part in the max (utils.py:27-29)
            else:
                denom = int(counts[1:].max())  # reference: max over non(-1) keys; raises when none are present
                if denom == 0:
                    raise ValueError("max() arg is an empty sequence")
            # python float division (double) then stored into an fp32 tensor element (utils.py:29)
            cls = counts[1:self.num_classes + 1].to(torch.float64) / float(denom)
            self.classwise_acc = cls.to(torch.float32)

    def masking(self, probs: Tensor, idx_ulb: Tensor, p_cutoff: float) -> Tensor:
        max_probs, max_idx = torch.max(probs, dim=-1)
        acc = self.classwise_acc[max_idx]
        self.last_gap = max_probs - p_cutoff * (acc / (2.0 - acc))   # diagnostic for the parity tests: distance to the threshold
        mask = max_probs.ge(p_cutoff * (acc / (2.0 - acc))).to(max_probs.dtype)  # utils.py:52
        select = max_probs.ge(p_cutoff)
        if int(select.sum()) != 0:
            self.selected_label[idx_ulb[select]] = max_idx[select]
        self.update()
        return mask


class FreeMatchState:
    """FreeMatchThresholdingHook (freematch/utils.py:10-66), world_size 1 view (probs already gathered)."""

    def __init__(self, num_classes: int, momentum: float = 0.999):
        self.m = momentum
        self.p_model = torch.ones(num_classes) / num_classes
        self.label_hist = torch.ones(num_classes) / num_classes
        self.time_p = self.p_model.mean()

    def update(self, probs: Tensor, use_quantile: bool, clip_thresh: bool):
        max_probs, max_idx = torch.max(probs, dim=-1, keepdim=True)
        if use_quantile:
            self.time_p = self.time_p * self.m + (1 - self.m) * torch.quantile(max_probs, 0.8)
        else:
            self.time_p = self.time_p * self.m + (1 - self.m) * max_probs.mean()
        if clip_thresh:
            self.time_p = torch.clip(self.time_p, 0.0, 0.95)
        self.p_model = self.p_model * self.m + (1 - self.m) * probs.mean(dim=0)
        hist = torch.bincount(max_idx.reshape(-1), minlength=self.p_model.shape[0]).to(self.p_model.dtype)
        self.label_hist = self.label_hist * self.m + (1 - self.m) * (hist / hist.sum())

    def masking(self, probs: Tensor, use_quantile: bool, clip_thresh: bool) -> Tensor:
        self.update(probs, use_quantile, clip_thresh)
        max_probs, max_idx = probs.max(dim=-1)
        mod = self.p_model / torch.max(self.p_model, dim=-1)[0]
        self.last_gap = max_probs - self.time_p * mod[max_idx]   # diagnostic for the parity tests: distance to the threshold
        return max_probs.ge(self.time_p * mod[max_idx]).to(max_probs.dtype)


class SoftMatchState:
    """SoftMatchWeightingHook, per_class=False (srsoftmatch/utils.py:12-77)."""

    def __init__(self, num_classes: int, n_sigma: int = 2, momentum: float = 0.999):
        self.m, self.n_sigma = momentum, n_sigma
        self.prob_max_mu_t = torch.tensor(1.0 / num_classes)
        self.prob_max_var_t = torch.tensor(1.0)

    def masking(self, probs: Tensor) -> Tensor:
        max_probs, _ = probs.max(dim=-1)
        mu = torch.mean(max_probs).item()  # .item(): python double enters the EMA (utils.py:39-40)
        var = torch.var(max_probs, unbiased=True).item()
        self.prob_max_mu_t = self.m * self.prob_max_mu_t + (1 - self.m) * mu
        self.prob_max_var_t = self.m * self.prob_max_var_t + (1 - self.m) * var
        return torch.exp(-((torch.clamp(max_probs - self.prob_max_mu_t, max=0.0) ** 2)
                           / (2 * self.prob_max_var_t / (self.n_sigma ** 2))))


class DistAlignState:
    """DistAlignEMAHook with p_target_type='uniform' (hooks/dist_align.py:10-72)."""

    def __init__(self, num_classes: int, momentum: float = 0.999):
        self.m = momentum
        self.p_target = torch.ones(num_classes) / num_classes
        self.p_model: Optional[Tensor] = None

    def dist_align(self, probs: Tensor) -> Tensor:
        mean = torch.mean(probs, dim=0)
        self.p_model = mean if self.p_model is None else self.p_model * self.m + mean * (1 - self.m)
        aligned = probs * (self.p_target + 1e-6) / (self.p_model + 1e-6)
        return aligned / aligned.sum(dim=-1, keepdim=True)


# ------------------------------------------------------------------------------------------------
# Optimizers / schedule
# ------------------------------------------------------------------------------------------------
def cosine_lr_factor(step: int, num_training_steps: int, num_warmup_steps: int = 0, num_cycles: float = 7.0 / 16.0) -> float:
    """build.py:227-251."""
    if step < num_warmup_steps:
        return float(step) / float(max(1, num_warmup_steps))
    s = float(step - num_warmup_steps) / float(max(1, num_training_steps - num_warmup_steps))
    return max(0.0, math.cos(math.pi * num_cycles * s))


def vit_layer_id(name: str, depth: int) -> int:
    """layer id of the reference's group_matcher + param_groups_layer_decay (vit.py:311-320, nets/utils.py:143-204):
    stem (cls_token,pos_embed,patch_embed) -> 0, blocks.i -> i+1, final norm joins the last block's group, head -> depth+1.
    Verified against the live reference's optimizer.param_groups (tests/test_oracle_golden.py (test_oracle_matches_live_reference))."""
    if name.startswith(("cls_token", "pos_embed", "patch_embed")):
        return 0
    if name.startswith("blocks."):
        return int(name.split(".")[1]) + 1
    if name.startswith("norm"):
        return depth
    return depth + 1


def vit_param_hparams(names_shapes, depth: int, lr: float, weight_decay: float, layer_decay: float):
    """{name: (lr_scale*lr, weight_decay)} as torch.optim.AdamW would see them (build.py:193-224)."""
    num_layers = depth + 2
    out = {}
    for n, shp in names_shapes:
        no_decay = len(shp) == 1 or n in ("pos_embed", "cls_token")
        scale = layer_decay ** (num_layers - 1 - vit_layer_id(n, depth)) if layer_decay != 1.0 else 1.0
        out[n] = (scale * lr, 0.0 if no_decay else weight_decay)
    return out


class AdamState:
    """torch.optim.Adam / AdamW (decoupled) single-tensor math, betas (0.9, 0.999), eps 1e-8, no amsgrad."""

    def __init__(self, params: Dict[str, Tensor], decoupled: bool):
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = {k: 0 for k in params}
        self.decoupled = decoupled

    @torch.no_grad()
    def step(self, params: Dict[str, Tensor], grads: Dict[str, Optional[Tensor]], lr_wd: Dict[str, Tuple[float, float]],
             b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
        for k, p in params.items():
            g = grads.get(k)
            if g is None:  # torch skips parameters whose .grad is None (Generator never gets one)
                continue
            lr, wd = lr_wd[k]
            self.t[k] += 1
            t = self.t[k]
            if wd != 0.0:
                if self.decoupled:
                    p.mul_(1 - lr * wd)
                else:
                    g = g.add(p, alpha=wd)
            self.m[k].lerp_(g, 1 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            bc1 = 1 - b1 ** t
            bc2 = 1 - b2 ** t
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(eps)
            p.addcdiv_(self.m[k], denom, value=-(lr / bc1))


# ------------------------------------------------------------------------------------------------
# The step
# ------------------------------------------------------------------------------------------------
@dataclass
class StepConfig:
    algorithm: str = "srflexmatch"  # srflexmatch | srfreematch | srsoftmatch | srfixmatch | srpseudolabel
    unsup_warm_up: float = 0.4  # srpseudolabel
    num_classes: int = 100
    ulb_dest_len: int = 50000
    p_cutoff: float = 0.95
    thresh_warmup: bool = True
    lambda_u: float = 1.0
    lambda_e: float = 0.001  # srfreematch ent_loss_ratio
    start_timing: int = 20000
    N_k: int = 10
    num_train_iter: int = 204800
    num_warmup_iter: int = 5120
    lr: float = 5e-4
    weight_decay: float = 5e-4
    layer_decay: float = 0.5
    sr_lr: float = 5e-4
    ema_p: float = 0.999
    use_quantile: bool = True
    clip_thresh: bool = False
    n_sigma: int = 2
    feature_dim: int = 384


def sr_decay(num_train_iter: int, it: int, max_sampling_time: int = 8) -> int:
    return int(max(max_sampling_time, 1 + num_train_iter / it))  # algorithmbase.py:182


class SSLOracle:
    """Holds backbone/Rewarder/Generator parameters, optimizer and hook state; `train_step` + `param_update`
    follow srflexmatch.py:107-217 / srfreematch.py:116-228 / srsoftmatch.py:107-221 and param_update.py:21-40."""

    def __init__(self, vit_cfg: Optional[ViTConfig], cfg: StepConfig, params: Dict[str, Tensor], rewarder: Dict[str, Tensor],
                 generator: Dict[str, Tensor], hparams: Optional[Dict[str, Tuple[float, float]]] = None):
        self.vit_cfg, self.cfg = vit_cfg, cfg
        self.p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        self.rp = {k: v.clone().requires_grad_(True) for k, v in rewarder.items()}
        self.gp = {k: v.clone().requires_grad_(True) for k, v in generator.items()}
        self.opt = AdamState(self.p, decoupled=True)
        self.ropt = AdamState(self.rp, decoupled=False)
        self.gopt = AdamState(self.gp, decoupled=False)
        # per-parameter (lr, weight decay); other backbones (oracle/bert_oracle.py) pass their own table
        self.hp = hparams if hparams is not None else vit_param_hparams(vit_cfg.param_shapes(), vit_cfg.depth, cfg.lr, cfg.weight_decay, cfg.layer_decay)
        self.sched_step = 0  # number of scheduler.step() calls so far (LambdaLR.last_epoch)
        self.max_reward = -float("inf")
        C = cfg.num_classes
        if cfg.algorithm in ("srfixmatch", "srpseudolabel"):
            self.hook = None  # FixedThresholdingHook is stateless (hooks/masking.py:42-57)
        elif cfg.algorithm == "srflexmatch":
            self.hook = FlexMatchState(cfg.ulb_dest_len, C, cfg.thresh_warmup)
        elif cfg.algorithm == "srfreematch":
            self.hook = FreeMatchState(C, cfg.ema_p)
        elif cfg.algorithm == "srsoftmatch":
            self.hook = SoftMatchState(C, cfg.n_sigma, cfg.ema_p)
            self.da = DistAlignState(C, cfg.ema_p)
        else:
            raise KeyError(f"Unknown algorithm: {cfg.algorithm}")
        self.loss: Optional[Tensor] = None
        self.drop_gen: Optional[torch.Generator] = None
        # Test-cost switch, OFF by default (the pinned behaviour is the reference's: K full backbone passes per stage-2 step).
        # With DropPath off the K passes are numerically identical, so a big-shape test may evaluate the backbone once per
        # data_generator call and replay only the hook-state updates — bit-equivalent in that deterministic mode ONLY.
        self.reuse_deterministic_passes = False
        # Data parallel: a torch.distributed group (gloo on CPU).  Every rank holds its own oracle on its own
        # shard; backbone gradients are averaged over the ranks (DDP, misc.py:56-64); of the SR update's two backward passes
        # only the first (generator_loss) is synchronised by DDP's reducer, the second (rewarder_loss) stays local
        # (srflexmatch.py:204-205; measured on the reference's own module by scripts/c3_ddp_probe.py).
        self.dp_group = None

    # -- pieces ---------------------------------------------------------------------------------
    def _backbone(self, x_lb, x_ulb_w, x_ulb_s):
        nb = x_lb.shape[0]
        x = torch.cat((x_lb, x_ulb_w, x_ulb_s))  # use_cat=True (srflexmatch.py:112-118)
        masks = draw_drop_path_masks(self.vit_cfg, x.shape[0], self.drop_gen)
        logits, feat = vit_forward(self.p, x, self.vit_cfg, masks)
        lw, ls = logits[nb:].chunk(2)
        fw, fs = feat[nb:].chunk(2)
        return logits[:nb], lw, ls, feat[:nb], fw, fs

    def _mask_from_probs(self, probs_w: Tensor, idx_ulb: Optional[Tensor]) -> Tensor:
        c = self.cfg
        if c.algorithm == "srfixmatch":  # FixedThresholdingHook.masking (hooks/masking.py:47-57; srfixmatch/fixmatch.py:132)
            return probs_w.max(dim=-1)[0].ge(c.p_cutoff).to(probs_w.dtype)
        if c.algorithm == "srflexmatch":
            return self.hook.masking(probs_w, idx_ulb, c.p_cutoff)
        if c.algorithm == "srfreematch":
            return self.hook.masking(probs_w, c.use_quantile, c.clip_thresh)
        return self.hook.masking(probs_w)

    def _data_generator(self, x_lb, idx_ulb, x_ulb_w, x_ulb_s, it: int, rec: dict):
        """data_generator (srflexmatch.py:72-104 and twins): K passes, the last one survives."""
        K = sr_decay(self.cfg.num_train_iter, it)
        rec["K"] = K
        unsup = None
        reuse = self.reuse_deterministic_passes and self.vit_cfg is not None and self.vit_cfg.drop_path_rate == 0.0
        cached = None
        for _ in range(K):
            if not reuse or cached is None:
                cached = self._backbone(x_lb, x_ulb_w, x_ulb_s)
            _, lw, ls, _, fw, _ = cached
            probs_w = torch.softmax(lw.detach(), dim=-1)
            pseudo = probs_w.argmax(dim=-1)
            mask = self._mask_from_probs(probs_w, idx_ulb)  # un-aligned softmax probs in all three algorithms
            reward = rewarder_forward(self.rp, fw, pseudo)  # rewarder.eval(): no dropout/BN inside -> same math
            mask2 = torch.where(reward >= reward.mean(), 1, 0).squeeze().float()
            unsup = consistency_loss(ls, pseudo, mask, mask2)
            rec.update(dg_mask=mask, dg_mask2=mask2, dg_reward=reward.detach(), dg_pseudo=pseudo, dg_logits_w=lw.detach(),
                       dg_gap=getattr(self.hook, "last_gap", None))
            rec.setdefault("dg_all_logits_w", []).append(lw.detach())
            if hasattr(self.hook, "time_p"):
                rec.setdefault("dg_all_time_p", []).append(float(self.hook.time_p))
        return unsup

    def _sr_update(self, feats: Tensor, true_labels: Tensor, rec: dict):
        """Stage-1 (labelled) and stage-2 (pseudo-labelled) Rewarder/Generator update — both do the same thing
        (srflexmatch.py:173-208)."""
        C = self.cfg.num_classes
        gen = generator_forward(self.gp, feats).long()  # .long() cuts the graph: the Generator gets no gradient
        reward = rewarder_forward(self.rp, feats, gen.squeeze(1))
        target = sr_target(gen.squeeze(1), true_labels, C)
        gen_loss = F.mse_loss(reward, torch.ones_like(reward))
        rew_loss = F.mse_loss(reward, target)
        names = list(self.rp.keys())
        g1 = torch.autograd.grad(gen_loss, [self.rp[k] for k in names], retain_graph=True, allow_unused=True)
        g2 = torch.autograd.grad(rew_loss, [self.rp[k] for k in names], retain_graph=True, allow_unused=True)
        grads = {}
        if self.dp_group is not None:
            g1 = [self._dp_mean(torch.zeros_like(self.rp[k]) if a is None else a) for k, a in zip(names, g1)]
        for k, a, b in zip(names, g1, g2):
            grads[k] = None if (a is None and b is None) else ((0 if a is None else a) + (0 if b is None else b))
        lr_wd = {k: (self.cfg.sr_lr, 0.0) for k in names}
        self.gopt.step(self.gp, {}, {k: (self.cfg.sr_lr, 0.0) for k in self.gp})  # all grads None -> no-op
        self.ropt.step(self.rp, grads, lr_wd)
        rec.update(sr_grads={k: (None if v is None else v.detach().clone()) for k, v in grads.items()},
                   sr_gen_label=gen.squeeze(1), sr_reward=reward.detach(), sr_target=target,
                   sr_gen_loss=gen_loss.detach(), sr_rew_loss=rew_loss.detach())

    def _train_step_pseudolabel(self, batch: Dict[str, Tensor], it: int) -> dict:
        """SRPseudoLabel.train_step / data_generator (srpseudolabel/srpseudolabel.py:58-201, task_type 'cls'): one unlabelled
        view whose own logits carry the gradient of the consistency loss; the model is called separately on x_lb and
        x_ulb_w (identical to a concatenated call for a LayerNorm net)."""
        c = self.cfg
        x_lb, y_lb, x_u = batch["x_lb"], batch["y_lb"], batch["x_ulb_w"]
        rec: dict = {}

        def fwd(x):
            masks = draw_drop_path_masks(self.vit_cfg, x.shape[0], self.drop_gen)
            return vit_forward(self.p, x, self.vit_cfg, masks)
        llb, flb = fwd(x_lb)
        lu, fu = fwd(x_u)
        sup_loss = ce_loss(llb, y_lb, "mean")
        probs = torch.softmax(lu.detach(), dim=-1)                       # FixedThresholdingHook with softmax_x_ulb=True (:105)
        mask = probs.max(dim=-1)[0].ge(c.p_cutoff).to(probs.dtype)
        pseudo = lu.detach().argmax(dim=-1)                               # gen_ulb_targets(logits, use_hard_label=True) (:108-110)
        if it > c.start_timing:
            K = sr_decay(c.num_train_iter, it)
            rec["K"] = K
            for _ in range(K):                                            # data_generator (:58-88)
                lk, fk = fwd(x_u)
                pk = torch.softmax(lk.detach(), dim=-1)
                mk = pk.max(dim=-1)[0].ge(c.p_cutoff).to(pk.dtype)
                pseudo_k = lk.detach().argmax(dim=-1)
                reward = rewarder_forward(self.rp, fk, pseudo_k)
                mask2 = torch.where(reward >= reward.mean(), 1, 0).squeeze().float()
                unsup_loss = consistency_loss(lk, pseudo_k, mk, mask2)
                rec.update(dg_mask=mk, dg_mask2=mask2, dg_reward=reward.detach(), dg_pseudo=pseudo_k)
        else:
            unsup_loss = consistency_loss(lu, pseudo, mask)
        if it > 0:
            if it >= c.start_timing:
                r = rewarder_forward(self.rp, fu.detach(), pseudo).mean()
                self.max_reward = float(r) if float(r) > self.max_reward else self.max_reward
                if it % c.N_k == 0 and it > c.start_timing:
                    self.max_reward = -float("inf")
                    self._sr_update(fu.detach(), pseudo, rec)
            else:
                self._sr_update(flb.detach(), y_lb, rec)
        warm = float(np.clip(it / (c.unsup_warm_up * c.num_train_iter), a_min=0.0, a_max=1.0))   # :194
        total = sup_loss + c.lambda_u * unsup_loss * warm
        self.loss = total
        rec.update(logits_lb=llb.detach(), logits_w=lu.detach(), feat_lb=flb.detach(), feat_w=fu.detach(), probs_w=probs, pseudo=pseudo,
                   mask=mask, sup_loss=sup_loss.detach(), unsup_loss=unsup_loss.detach(), total_loss=total.detach(),
                   util_ratio=mask.float().mean())
        return rec

    # -- public ---------------------------------------------------------------------------------
    def train_step(self, batch: Dict[str, Tensor], it: int) -> dict:
        c = self.cfg
        if c.algorithm == "srpseudolabel":
            return self._train_step_pseudolabel(batch, it)
        x_lb, y_lb, x_ulb_w, x_ulb_s = batch["x_lb"], batch["y_lb"], batch["x_ulb_w"], batch["x_ulb_s"]
        idx_ulb = batch.get("idx_ulb")
        rec: dict = {}
        llb, lw, ls, flb, fw, fs = self._backbone(x_lb, x_ulb_w, x_ulb_s)
        sup_loss = ce_loss(llb, y_lb, "mean")
        probs_w = torch.softmax(lw.detach(), dim=-1)
        if c.algorithm == "srsoftmatch":
            probs_for_mask = self.da.dist_align(probs_w)  # srsoftmatch.py:138
        else:
            probs_for_mask = probs_w
        mask = self._mask_from_probs(probs_for_mask, idx_ulb)
        rec["gap"] = getattr(self.hook, "last_gap", None)
        pseudo = lw.detach().argmax(dim=-1)  # argmax(probs) == argmax(logits) up to fp ties; reference uses probs for
        if c.algorithm in ("srflexmatch", "srfixmatch"):     # FlexMatch / FixMatch (srflexmatch.py:142-146, fixmatch.py:135-139) and logits for the others
            pseudo = probs_w.argmax(dim=-1)
        if it > c.start_timing:
            unsup_loss = self._data_generator(x_lb, idx_ulb, x_ulb_w, x_ulb_s, it, rec)
        else:
            unsup_loss = consistency_loss(ls, pseudo, mask)
        if it > 0:
            if it >= c.start_timing:
                r = rewarder_forward(self.rp, fw.detach(), pseudo).mean()
                self.max_reward = float(r) if float(r) > self.max_reward else self.max_reward
                rec["sr_mean_reward"] = r.detach()
                # the where() at srflexmatch.py:171-172 is always False -> "filtered" == current batch
                if it % c.N_k == 0 and it > c.start_timing:
                    self.max_reward = -float("inf")
                    self._sr_update(fw.detach(), pseudo, rec)
            else:
                self._sr_update(flb.detach(), y_lb, rec)
        total = sup_loss + c.lambda_u * unsup_loss
        if c.algorithm == "srfreematch":
            if float(mask.sum()) > 0:
                total = total + c.lambda_e * freematch_entropy_loss(mask, ls, self.hook.p_model, self.hook.label_hist)
        self.loss = total
        rec.update(logits_lb=llb.detach(), logits_w=lw.detach(), logits_s=ls.detach(), feat_lb=flb.detach(),
                   feat_w=fw.detach(), feat_s=fs.detach(), probs_w=probs_w, pseudo=pseudo, mask=mask,
                   sup_loss=sup_loss.detach(), unsup_loss=unsup_loss.detach(), total_loss=total.detach(),
                   util_ratio=mask.float().mean())
        return rec

    def _dp_mean(self, t: Tensor) -> Tensor:
        import torch.distributed as dist
        t = t.detach().clone()
        dist.all_reduce(t, group=self.dp_group)
        return t / dist.get_world_size(self.dp_group)

    def param_update(self) -> Dict[str, Tensor]:
        """ParamUpdateHook.after_train_step (param_update.py:21-40): backward, AdamW step, scheduler step, zero_grad.
        Returns the gradients (for parity checks)."""
        names = list(self.p.keys())
        gs = torch.autograd.grad(self.loss, [self.p[k] for k in names], allow_unused=True)
        grads = {k: g for k, g in zip(names, gs)}
        if self.dp_group is not None:
            grads = {k: self._dp_mean(torch.zeros_like(self.p[k]) if g is None else g) for k, g in grads.items()}
        f = cosine_lr_factor(self.sched_step, self.cfg.num_train_iter, self.cfg.num_warmup_iter)
        lr_wd = {k: (self.hp[k][0] * f, self.hp[k][1]) for k in names}
        self.opt.step(self.p, grads, lr_wd)
        self.sched_step += 1
        self.loss = None
        return grads


def build_det_oracle(vit_cfg: ViTConfig, cfg: StepConfig, seed: int = 0, head_gain: float = 1.0) -> SSLOracle:
    """Oracle initialised from semireward_b200.detgen fills (same tensors the golden generator loads into the
    reference)."""
    from semireward_b200 import detgen
    p = {n: torch.from_numpy(detgen.fill_param(n, s, seed)) for n, s in vit_cfg.param_shapes()}
    if head_gain != 1.0:
        p["head.weight"] = p["head.weight"] * head_gain
    rp = {n: torch.from_numpy(detgen.fill_param("rewarder." + n, s, seed))
          for n, s in rewarder_param_shapes(cfg.feature_dim, cfg.num_classes)}
    gp = {n: torch.from_numpy(detgen.fill_param("generator." + n, s, seed)) for n, s in generator_param_shapes(cfg.feature_dim)}
    return SSLOracle(vit_cfg, cfg, p, rp, gp)


def to_torch_batch(batch: Dict[str, np.ndarray]) -> Dict[str, Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in batch.items()}
