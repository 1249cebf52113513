"""Optimizer + schedule with the reference's construction API (semilearn/core/utils/build.py:193-251,
semilearn/nets/utils.py:77-204): `get_optimizer(net, 'AdamW', lr, momentum, weight_decay, layer_decay)` builds the same
param groups (layer-wise lr decay through the net's group_matcher, no weight decay on 1-D tensors and on
net.no_weight_decay()), `get_cosine_schedule_with_warmup` returns the same LambdaLR.  The optimizer itself is
FusedAdamW: a torch.optim.Optimizer whose step() is ONE srw_adamw_step launch over all tensors, which also rewrites the
ViT engine's split-bf16 weight cache in the same pass."""
from __future__ import annotations

import ctypes as C
import math
import re

import numpy as np
import torch

from .. import _lib as L


def _layer_id(name: str, matcher: dict) -> int | None:
    """Layer index of a parameter from the net's group_matcher (vit.py:311-320): stem -> 0, blocks.i -> i+1,
    patterns tagged (99999,) (the final norm) join the last block group, everything unmatched (head) -> last+1."""
    if re.match(matcher["stem"], name):
        return 0
    blocks = matcher["blocks"]
    if isinstance(blocks, str):   # bert.py:58-60 / hubert.py:51-53: one pattern whose group is the layer index
        m = re.match(blocks, name)
        return int(m.group(1)) + 1 if m else None
    for pat, tag in blocks:
        m = re.match(pat, name)
        if m:
            return ("tail" if tag is not None else int(m.group(1)) + 1)
    return None


def param_groups_layer_decay(net, lr, weight_decay=0.05, no_weight_decay_list=(), layer_decay=0.75):
    """Same grouping as the reference (nets/utils.py:143-204): one group per (layer id, decay / no_decay) with
    lr_scale = layer_decay ** (num_layers - 1 - layer_id) and 'lr' pre-multiplied."""
    matcher = net.group_matcher(coarse=False)
    named = [(n, p) for n, p in net.named_parameters() if p.requires_grad]
    ids = {n: _layer_id(n, matcher) for n, _ in named}
    max_block = max([v for v in ids.values() if isinstance(v, int)] + [0])
    resolved = {}
    for n, v in ids.items():
        resolved[n] = max_block if v == "tail" else (max_block + 1 if v is None else v)
    num_layers = max_block + 2
    groups: dict = {}
    for n, p in named:
        no_decay = p.ndim == 1 or n in no_weight_decay_list
        lid = resolved[n]
        key = (lid, no_decay)
        if key not in groups:
            scale = layer_decay ** (num_layers - 1 - lid)
            groups[key] = dict(lr_scale=scale, lr=scale * lr, weight_decay=0.0 if no_decay else weight_decay, params=[], param_names=[])
        groups[key]["params"].append(p)
        groups[key]["param_names"].append(n)
    return list(groups.values())


def param_groups_weight_decay(net, weight_decay=1e-5, no_weight_decay_list=()):
    decay, no_decay = [], []
    for n, p in net.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if (p.ndim <= 1 or n.endswith(".bias") or n in no_weight_decay_list) else decay).append(p)
    return [dict(params=no_decay, weight_decay=0.0), dict(params=decay, weight_decay=weight_decay)]


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW-compatible (param_groups, state_dict keys 'step' / 'exp_avg' / 'exp_avg_sq') but stepping all
    tensors with one native launch.  `net` (optional): a semireward_b200 ViT whose weight-plane cache is refreshed
    inside the same kernel."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, net=None, decoupled=True):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._net = net
        self._decoupled = decoupled
        self._rows = None
        self._dev_table = None
        self._host_tables = None
        self._flip = 0
        self._step = 0

    def load_state_dict(self, state_dict):
        """torch's loader replaces the per-parameter state tensors; the flat moment buffers and the row table are rebuilt
        from them on the next step() (`_build` copies a restored exp_avg / exp_avg_sq into the flat buffers)."""
        super().load_state_dict(state_dict)
        self._rows = None
        self._step = 0

    def _build(self):
        g0 = self.param_groups[0]
        for g in self.param_groups:   # one launch = one (betas, eps) pair; the reference's groups only differ in lr / weight decay
            if tuple(g["betas"]) != tuple(g0["betas"]) or g["eps"] != g0["eps"]:
                raise ValueError("FusedAdamW: all param groups must share betas and eps")
        # torch.optim.AdamW skips parameters whose .grad is None (no weight decay, no moments): BERT's pooler, whose output the
        # wrapper never uses (bert.py:35).  The table is built at the first step, over the parameters that have a gradient then.
        ps = [p for g in self.param_groups for p in g["params"] if p.grad is not None]
        self._skipped = sum(1 for g in self.param_groups for p in g["params"] if p.grad is None)
        dev = ps[0].device
        total = sum(p.numel() for p in ps)
        self._m_flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self._v_flat = torch.zeros(total, dtype=torch.float32, device=dev)
        slot = {}
        if self._net is not None and hasattr(self._net, "_ordered_params"):
            planes = self._net._weight_planes()  # allocates + fills the cache once
            lib = L.load()
            for idx, p in enumerate(self._net._ordered_params()):
                if hasattr(self._net, "weight_plane_slot"):
                    sl = self._net.weight_plane_slot(idx)
                    if sl is not None:
                        slot[id(p)] = (planes.data_ptr() + sl[0],) + tuple(sl[1:])
                    continue
                off, cols, ldp, ps_ = L.i64(), L.i32(), L.i32(), L.i64()
                if lib.srw_vit_weight_plane_slot(C.byref(self._net._cfg), idx, C.byref(off), C.byref(cols), C.byref(ldp), C.byref(ps_)) == 0:
                    slot[id(p)] = (planes.data_ptr() + off.value, cols.value, ldp.value, ps_.value)
        rows = (L.AdamWRow * len(ps))()
        off = blk = 0
        self._group_of = []
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                if p.grad is None:
                    continue
                r = rows[len(self._group_of)]
                n = p.numel()
                m, v = self._m_flat[off:off + n].view_as(p), self._v_flat[off:off + n].view_as(p)
                st = self.state[p]
                if "exp_avg" in st:   # state restored by load_state_dict
                    m.copy_(st["exp_avg"]); v.copy_(st["exp_avg_sq"])
                st["exp_avg"], st["exp_avg_sq"] = m, v
                st.setdefault("step", torch.tensor(0.0))
                r.param, r.exp_avg, r.exp_avg_sq, r.numel = p.data_ptr(), m.data_ptr(), v.data_ptr(), n
                pl = slot.get(id(p))
                if pl is not None:
                    r.planes, r.cols, r.ldp, r.plane_stride = pl
                else:
                    r.planes, r.cols, r.ldp, r.plane_stride = None, 1, 1, 0
                r.first_block = blk
                blk += (n + L.ADAMW_BLOCK_ELEMS - 1) // L.ADAMW_BLOCK_ELEMS
                off += n
                self._group_of.append(gi)
        self._rows, self._params, self._total_blocks = rows, ps, blk
        nbytes = C.sizeof(rows)
        self._dev_table = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self._host_tables = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._copy_events = [None, None]
        if self.state and any(float(s.get("step", 0)) > 0 for s in self.state.values()):
            self._step = int(max(float(s["step"]) for s in self.state.values()))

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FusedAdamW.step does not take a closure")
        if self._rows is None or sum(1 for g in self.param_groups for p in g["params"] if p.grad is None) != self._skipped:
            self._build()
        rows = self._rows
        for i, p in enumerate(self._params):
            g = p.grad
            if g is None:
                raise RuntimeError("FusedAdamW: a parameter lost its gradient between steps")
            if not g.is_contiguous():
                g = p.grad = g.contiguous()
            grp = self.param_groups[self._group_of[i]]
            rows[i].grad, rows[i].lr, rows[i].weight_decay = g.data_ptr(), float(grp["lr"]), float(grp["weight_decay"])
        self._step += 1
        k = self._flip
        self._flip ^= 1
        if self._copy_events[k] is not None:
            self._copy_events[k].synchronize()   # the staging buffer's previous H2D copy must have been consumed
        host = self._host_tables[k]
        C.memmove(host.data_ptr(), C.addressof(rows), C.sizeof(rows))
        self._dev_table.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._copy_events[k] = ev
        g0 = self.param_groups[0]
        a = L.AdamWArgs(num_tensors=len(self._params), total_blocks=self._total_blocks, table=self._dev_table.data_ptr(), lr_factor=1.0,
                        beta1=g0["betas"][0], beta2=g0["betas"][1], eps=g0["eps"], step=self._step, decoupled=int(self._decoupled))
        L.check(L.load().srw_adamw_step(C.byref(a), L.stream_ptr()), "srw_adamw_step")
        if self._net is not None and hasattr(self._net, "mark_weights_updated"):
            self._net.mark_weights_updated(planes_fresh=True)
        for s in self.state.values():
            if "step" in s:
                s["step"] = torch.tensor(float(self._step))
        return None


class FusedSGD(torch.optim.Optimizer):
    """torch.optim.SGD-compatible (param_groups, state_dict key 'momentum_buffer') stepping all tensors with one srw_sgd_step launch
    (the optimizer of config/classic_cv: SGD lr 0.03, momentum 0.9, nesterov, core/utils/build.py:219-220).  Parameters whose
    .grad is None are skipped like torch does (WideResNet's two unused bn1)."""

    def __init__(self, params, lr=0.03, momentum=0.9, weight_decay=0.0, nesterov=True, net=None):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, nesterov=nesterov, dampening=0))
        self._net = net
        self._rows = None
        self._first = True
        self._flip = 0

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._rows = None

    def _build(self):
        g0 = self.param_groups[0]
        for g in self.param_groups:
            if g["momentum"] != g0["momentum"] or g["nesterov"] != g0["nesterov"] or g.get("dampening", 0) != 0:
                raise ValueError("FusedSGD: all param groups must share momentum / nesterov, dampening must be 0")
        ps = [p for g in self.param_groups for p in g["params"] if p.grad is not None]
        self._skipped = sum(1 for g in self.param_groups for p in g["params"] if p.grad is None)
        dev = ps[0].device
        self._buf_flat = torch.zeros(sum(p.numel() for p in ps), dtype=torch.float32, device=dev)
        rows = (L.AdamWRow * len(ps))()
        off = blk = 0
        self._group_of = []
        restored = False
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                if p.grad is None:
                    continue
                r = rows[len(self._group_of)]
                n = p.numel()
                m = self._buf_flat[off:off + n].view_as(p)
                st = self.state[p]
                if st.get("momentum_buffer") is not None:   # restored by load_state_dict
                    m.copy_(st["momentum_buffer"])
                    restored = True
                st["momentum_buffer"] = m
                r.param, r.exp_avg, r.exp_avg_sq, r.numel = p.data_ptr(), m.data_ptr(), None, n
                r.planes, r.cols, r.ldp, r.plane_stride = None, 1, 1, 0
                r.first_block = blk
                blk += (n + L.ADAMW_BLOCK_ELEMS - 1) // L.ADAMW_BLOCK_ELEMS
                off += n
                self._group_of.append(gi)
        self._rows, self._params, self._total_blocks = rows, ps, blk
        if restored:
            self._first = False
        nbytes = C.sizeof(rows)
        self._dev_table = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self._host_tables = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._copy_events = [None, None]

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FusedSGD.step does not take a closure")
        if self._rows is None or sum(1 for g in self.param_groups for p in g["params"] if p.grad is None) != self._skipped:
            self._build()
        rows = self._rows
        for i, p in enumerate(self._params):
            g = p.grad
            if g is None:
                raise RuntimeError("FusedSGD: a parameter lost its gradient between steps")
            if not g.is_contiguous():
                g = p.grad = g.contiguous()
            grp = self.param_groups[self._group_of[i]]
            rows[i].grad, rows[i].lr, rows[i].weight_decay = g.data_ptr(), float(grp["lr"]), float(grp["weight_decay"])
        k = self._flip
        self._flip ^= 1
        if self._copy_events[k] is not None:
            self._copy_events[k].synchronize()
        host = self._host_tables[k]
        C.memmove(host.data_ptr(), C.addressof(rows), C.sizeof(rows))
        self._dev_table.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._copy_events[k] = ev
        g0 = self.param_groups[0]
        a = L.SgdArgs(num_tensors=len(self._params), total_blocks=self._total_blocks, table=self._dev_table.data_ptr(), lr_factor=1.0,
                      momentum=float(g0["momentum"]), nesterov=int(bool(g0["nesterov"])), first_step=int(self._first))
        L.check(L.load().srw_sgd_step(C.byref(a), L.stream_ptr()), "srw_sgd_step")
        self._first = False
        if self._net is not None and hasattr(self._net, "mark_weights_updated"):
            self._net.mark_weights_updated()      # the kernel wrote the parameters behind torch's version counters
        return None


def get_optimizer(net, optim_name="SGD", lr=0.1, momentum=0.9, weight_decay=0, layer_decay=1.0, nesterov=True, bn_wd_skip=True):
    assert layer_decay <= 1.0
    no_decay = net.no_weight_decay() if (hasattr(net, "no_weight_decay") and bn_wd_skip) else {}
    if layer_decay != 1.0:
        groups = param_groups_layer_decay(net, lr, weight_decay, no_weight_decay_list=no_decay, layer_decay=layer_decay)
    else:
        groups = param_groups_weight_decay(net, weight_decay, no_weight_decay_list=no_decay)
    if optim_name == "AdamW":
        return FusedAdamW(groups, lr=lr, weight_decay=weight_decay, net=net)
    if optim_name == "SGD":
        if hasattr(net, "forward_native"):     # a native backbone: one fused launch; anything else keeps torch's optimizer
            return FusedSGD(groups, lr=lr, momentum=momentum, weight_decay=weight_decay, nesterov=nesterov, net=net)
        return torch.optim.SGD(groups, lr=lr, momentum=momentum, weight_decay=weight_decay, nesterov=nesterov)
    raise ValueError(f"unknown optimizer {optim_name}")


def get_cosine_schedule_with_warmup(optimizer, num_training_steps, num_cycles=7.0 / 16.0, num_warmup_steps=0, last_epoch=-1):
    from torch.optim.lr_scheduler import LambdaLR

    def _lr_lambda(current_step):
        if current_step < num_warmup_steps:
            return float(current_step) / float(max(1, num_warmup_steps))
        s = float(current_step - num_warmup_steps) / float(max(1, num_training_steps - num_warmup_steps))
        return max(0.0, math.cos(math.pi * num_cycles * s))

    return LambdaLR(optimizer, _lr_lambda, last_epoch)
