"""TEST INFRASTRUCTURE — CPU restatement of the convolutional path of the hot loop (BASELINE
configs[0]: FlexMatch+SemiReward, WRN-28-2, CIFAR-100, SGD): `WideResNet` (semilearn/nets/wrn/wrn.py:30-146) under the
SSL step of oracle/ssl_oracle.py with `use_cat: True` and torch.optim.SGD (core/utils/build.py:193-224).

Only tests/ (and, later, smoke()/bench.py's CPU arm) may import this module; the product never does.  Pinned against
the live reference running on CPU in this container: tests/test_wrn_oracle.py, fixture tests/golden/wrn_*.npz made by
tests/golden/make_golden_wrn.py.

What makes this path different from the LayerNorm nets (and what a native build has to honour):
  * BatchNorm couples the rows of the concatenated batch (x_lb, x_ulb_w, x_ulb_s): the weak rows DO carry gradient through
    the batch statistics, so the backward runs over all 3 blocks of rows (9·B·F, not 7·B·F).
  * Every forward in train mode advances the running statistics with momentum 0.001 (wrn.py:11,33,37,97), including the K
    sampling passes of stage 2 and — a quirk of wrn.py:46-54 — the bn1 of the first layer of block2 / block3, whose output
    is computed and then NOT used (`conv1(out if self.equalInOut else x)` takes the raw x when the channel count changes
    and activate_before_residual is False).  Those two bn1 get no gradient, so SGD never touches their weight / bias.
  * The final BatchNorm uses eps 1e-3, the others 1e-5 (wrn.py:97 vs the nn.BatchNorm2d default).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import ssl_oracle as O

Tensor = torch.Tensor


@dataclass
class WRNCfg:
    depth: int = 28
    widen: int = 2
    num_classes: int = 100
    first_stride: int = 1
    bn_momentum: float = 0.001
    slope: float = 0.1                 # LeakyReLU negative slope (wrn.py:34,38,98)

    @property
    def n(self) -> int:
        return (self.depth - 4) // 6

    @property
    def channels(self) -> List[int]:
        return [16, 16 * self.widen, 32 * self.widen, 64 * self.widen]

    def blocks(self):
        """(prefix, in_planes, out_planes, stride, activate_before_residual) of every BasicBlock, in module order."""
        ch = self.channels
        out = []
        for b, (cin, cout, stride, abr) in enumerate(((ch[0], ch[1], self.first_stride, True), (ch[1], ch[2], 2, False), (ch[2], ch[3], 2, False))):
            for i in range(self.n):
                out.append((f"block{b + 1}.layer.{i}.", cin if i == 0 else cout, cout, stride if i == 0 else 1, abr))
        return out

    def param_shapes(self) -> List[Tuple[str, Tuple[int, ...]]]:
        """`WideResNet.named_parameters()` order (81 tensors, 1 479 236 parameters for WRN-28-2 with 100 classes)."""
        ch = self.channels
        out = [("conv1.weight", (ch[0], 3, 3, 3)), ("conv1.bias", (ch[0],))]
        for pre, cin, cout, _, _ in self.blocks():
            out += [(pre + "bn1.weight", (cin,)), (pre + "bn1.bias", (cin,)), (pre + "conv1.weight", (cout, cin, 3, 3)),
                    (pre + "bn2.weight", (cout,)), (pre + "bn2.bias", (cout,)), (pre + "conv2.weight", (cout, cout, 3, 3))]
            if cin != cout:
                out.append((pre + "convShortcut.weight", (cout, cin, 1, 1)))
        out += [("bn1.weight", (ch[3],)), ("bn1.bias", (ch[3],)), ("classifier.weight", (self.num_classes, ch[3])),
                ("classifier.bias", (self.num_classes,))]
        return out

    def bn_names(self) -> List[Tuple[str, int]]:
        out = []
        for pre, cin, cout, _, _ in self.blocks():
            out += [(pre + "bn1", cin), (pre + "bn2", cout)]
        return out + [("bn1", self.channels[3])]

    def fwd_flops_per_image(self, img: int = 32) -> float:
        """2*MACs of the convolutions + classifier for one img x img image (0.429 GF... the probe counted
        multiply-adds of conv + linear only)."""
        hw = img * img
        fl = 2.0 * hw * 27 * self.channels[0]
        for _, cin, cout, stride, _ in self.blocks():
            hw_out = hw // (stride * stride)
            fl += 2.0 * hw_out * 9 * cin * cout + 2.0 * hw_out * 9 * cout * cout
            if cin != cout:
                fl += 2.0 * hw_out * cin * cout
            hw = hw_out
        return fl + 2.0 * self.channels[3] * self.num_classes


def new_bn_buffers(cfg: WRNCfg) -> Dict[str, Tensor]:
    buf = {}
    for name, c in cfg.bn_names():
        buf[name + ".running_mean"] = torch.zeros(c)
        buf[name + ".running_var"] = torch.ones(c)
    return buf


def _sync_batch_norm(t: Tensor, rm: Tensor, rv: Tensor, w: Tensor, b: Tensor, momentum: float, eps: float, group) -> Tensor:
    """nn.SyncBatchNorm in train mode (what send_model_cuda turns every BatchNorm2d into under DDP, core/utils/misc.py:54):
    mean / biased variance over the rows of ALL ranks, running statistics advanced with the global mean and the unbiased
    global variance.  Differentiable all-reduces, so the backward exchanges the gradient statistics as SyncBatchNorm does."""
    import torch.distributed.nn.functional as dfn
    C = t.shape[1]
    cnt = torch.tensor([float(t.numel() // C)])
    s = dfn.all_reduce(t.sum((0, 2, 3)), group=group)
    ss = dfn.all_reduce((t * t).sum((0, 2, 3)), group=group)
    n = dfn.all_reduce(cnt, group=group)
    mean = s / n
    var = ss / n - mean * mean
    with torch.no_grad():
        rm.mul_(1 - momentum).add_(mean.detach(), alpha=momentum)
        rv.mul_(1 - momentum).add_(var.detach() * (n / (n - 1)), alpha=momentum)
    return (t - mean[None, :, None, None]) * torch.rsqrt(var + eps)[None, :, None, None] * w[None, :, None, None] + b[None, :, None, None]


def wrn_forward(p: Dict[str, Tensor], buf: Dict[str, Tensor], x: Tensor, cfg: WRNCfg, training: bool = True, sync_group=None):
    """-> (logits [B, C], feat [B, 64*widen]).  In training mode the batch statistics of ALL rows of x normalise every row
    and `buf` (running mean / unbiased running var) is advanced in place, exactly like nn.BatchNorm2d.  `sync_group`: a
    torch.distributed group -> SyncBatchNorm semantics (statistics over the rows of every rank), the data-parallel form."""
    def bn(name, t, eps=1e-5):
        if sync_group is not None and training:
            return _sync_batch_norm(t, buf[name + ".running_mean"], buf[name + ".running_var"], p[name + ".weight"], p[name + ".bias"],
                                    cfg.bn_momentum, eps, sync_group)
        return F.batch_norm(t, buf[name + ".running_mean"], buf[name + ".running_var"], p[name + ".weight"], p[name + ".bias"], training,
                            cfg.bn_momentum, eps)

    def act(t):
        return F.leaky_relu(t, cfg.slope)

    out = F.conv2d(x, p["conv1.weight"], p["conv1.bias"], stride=1, padding=1)
    for pre, cin, cout, stride, abr in cfg.blocks():
        equal = cin == cout
        xin = out
        if not equal and abr:
            xin = act(bn(pre + "bn1", xin))            # replaces x: the shortcut sees the activated tensor too (wrn.py:47-48)
            o = xin
        else:
            o = act(bn(pre + "bn1", xin))              # computed even when unused below (running stats still advance)
            if not equal:
                o = xin                                 # wrn.py:51: conv1 takes the raw x when the channel count changes
        o = act(bn(pre + "bn2", F.conv2d(o, p[pre + "conv1.weight"], None, stride=stride, padding=1)))
        o = F.conv2d(o, p[pre + "conv2.weight"], None, stride=1, padding=1)
        sc = xin if equal else F.conv2d(xin, p[pre + "convShortcut.weight"], None, stride=stride, padding=0)
        out = torch.add(sc, o)
    out = act(bn("bn1", out, eps=1e-3))
    feat = F.adaptive_avg_pool2d(out, 1).view(-1, cfg.channels[3])
    return F.linear(feat, p["classifier.weight"], p["classifier.bias"]), feat


def wrn_param_hparams(names_shapes, lr: float, weight_decay: float):
    """param_groups_weight_decay (nets/utils.py:75-96) with WideResNet.no_weight_decay() (wrn.py:152-157): 1-D tensors, biases
    and everything with 'bn' in its name are not decayed; one learning rate (layer_decay 1.0 in configs[0])."""
    return {n: (lr, 0.0 if (len(s) <= 1 or n.endswith(".bias") or "bn" in n) else weight_decay) for n, s in names_shapes}


class SGDState:
    """torch.optim.SGD single-tensor math with momentum, dampening 0, nesterov (build.py:219-220)."""

    def __init__(self, params: Dict[str, Tensor], momentum: float = 0.9, nesterov: bool = True):
        self.buf: Dict[str, Optional[Tensor]] = {k: None for k in params}
        self.momentum, self.nesterov = momentum, nesterov

    @torch.no_grad()
    def step(self, params, grads, lr_wd):
        for k, p in params.items():
            g = grads.get(k)
            if g is None:     # torch skips parameters whose .grad is None (the two unused bn1, see the module docstring)
                continue
            lr, wd = lr_wd[k]
            if wd != 0.0:
                g = g.add(p, alpha=wd)
            if self.momentum != 0.0:
                if self.buf[k] is None:
                    self.buf[k] = g.clone()
                else:
                    self.buf[k].mul_(self.momentum).add_(g)
                g = g.add(self.buf[k], alpha=self.momentum) if self.nesterov else self.buf[k]
            p.add_(g, alpha=-lr)


class WRNSSLOracle(O.SSLOracle):
    """ssl_oracle.SSLOracle with the WRN backbone: one concatenated forward per pass (use_cat True), BatchNorm buffers as
    state, SGD instead of AdamW."""

    def __init__(self, wrn_cfg: WRNCfg, cfg: O.StepConfig, params, rewarder, generator, momentum: float = 0.9, nesterov: bool = True):
        hp = wrn_param_hparams(wrn_cfg.param_shapes(), cfg.lr, cfg.weight_decay)
        super().__init__(None, cfg, params, rewarder, generator, hparams=hp)
        self.wrn_cfg = wrn_cfg
        self.buf = new_bn_buffers(wrn_cfg)
        self.sgd = SGDState(self.p, momentum, nesterov)

    def _backbone(self, x_lb, x_ulb_w, x_ulb_s):
        nb = x_lb.shape[0]
        # data parallel: send_model_cuda converts the BatchNorms to SyncBatchNorm (misc.py:54) -> statistics over every rank's rows
        logits, feat = wrn_forward(self.p, self.buf, torch.cat((x_lb, x_ulb_w, x_ulb_s)), self.wrn_cfg, training=True,
                                   sync_group=getattr(self, "dp_group", None))
        lw, ls = logits[nb:].chunk(2)
        fw, fs = feat[nb:].chunk(2)
        return logits[:nb], lw, ls, feat[:nb], fw, fs

    def param_update(self):
        names = list(self.p.keys())
        gs = torch.autograd.grad(self.loss, [self.p[k] for k in names], allow_unused=True)
        grads = {k: g for k, g in zip(names, gs)}
        if getattr(self, "dp_group", None) is not None:   # DDP: average over the ranks (unused parameters stay without a gradient)
            grads = {k: (None if g is None else self._dp_mean(g)) for k, g in grads.items()}
        f = O.cosine_lr_factor(self.sched_step, self.cfg.num_train_iter, self.cfg.num_warmup_iter)
        self.sgd.step(self.p, grads, {k: (self.hp[k][0] * f, self.hp[k][1]) for k in names})
        self.sched_step += 1
        self.loss = None
        return grads


def build_det_wrn_oracle(wrn_cfg: WRNCfg, cfg: O.StepConfig, seed: int = 0, head_gain: float = 1.0) -> WRNSSLOracle:
    from semireward_b200 import detgen
    p = {n: torch.from_numpy(detgen.fill_param(n, s, seed)) for n, s in wrn_cfg.param_shapes()}
    if head_gain != 1.0:
        p["classifier.weight"] = p["classifier.weight"] * head_gain
    rp = {n: torch.from_numpy(detgen.fill_param("rewarder." + n, s, seed)) for n, s in O.rewarder_param_shapes(cfg.feature_dim, cfg.num_classes)}
    gp = {n: torch.from_numpy(detgen.fill_param("generator." + n, s, seed)) for n, s in O.generator_param_shapes(cfg.feature_dim)}
    return WRNSSLOracle(wrn_cfg, cfg, p, rp, gp)
