"""Hooks on the hot path, with the reference's names and call protocol (semilearn/core/hooks/hook.py:6-42):
`algorithm.call_hook(fn_name, hook_name, **kw)`.

  ParamUpdateHook            <- semilearn/core/hooks/param_update.py:21-45
  PseudoLabelingHook         <- semilearn/algorithms/hooks/pseudo_label.py:16-52
  FlexMatchThresholdingHook  <- semilearn/algorithms/srflexmatch/utils.py:11-63  (state lives on the device; the
                                softmax/argmax/mask/scatter/histogram run as ONE kernel, srw_flexmatch_mask)
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib as L


class Hook:
    stages = ("before_run", "before_train_epoch", "before_train_step", "after_train_step", "after_train_epoch", "after_run")

    def before_train_epoch(self, algorithm): pass
    def after_train_epoch(self, algorithm): pass
    def before_train_step(self, algorithm): pass
    def after_train_step(self, algorithm): pass
    def before_run(self, algorithm): pass
    def after_run(self, algorithm): pass

    # schedule helpers of the reference's Hook base (core/hooks/hook.py:28-42)
    def every_n_epochs(self, algorithm, n): return (algorithm.epoch + 1) % n == 0 if n > 0 else False
    def every_n_iters(self, algorithm, n): return (algorithm.it + 1) % n == 0 if n > 0 else False
    def is_last_epoch(self, algorithm): return algorithm.epoch + 1 == algorithm.epochs
    def is_last_iter(self, algorithm): return algorithm.it + 1 == algorithm.num_train_iter


class ParamUpdateHook(Hook):
    """backward + (clip) + optimizer.step + scheduler.step + zero_grad; CUDA-event run_time like the reference."""

    def before_train_step(self, algorithm):
        if hasattr(algorithm, "start_run"):
            torch.cuda.synchronize()
            algorithm.start_run.record()

    def after_train_step(self, algorithm):
        loss = algorithm.out_dict["loss"]
        if algorithm.use_amp:
            raise NotImplementedError("amp: every SemiReward config runs fp32 (amp: False); the native path is fp32-accurate")
        loss.backward()
        if algorithm.clip_grad > 0:
            torch.nn.utils.clip_grad_norm_(algorithm.model.parameters(), algorithm.clip_grad)
        algorithm.optimizer.step()
        if algorithm.scheduler is not None:
            algorithm.scheduler.step()
        algorithm.model.zero_grad()
        if hasattr(algorithm, "end_run"):
            algorithm.end_run.record()
            torch.cuda.synchronize()
            algorithm.log_dict["train/run_time"] = algorithm.start_run.elapsed_time(algorithm.end_run) / 1000.0


class EMA:
    """EMA of the model parameters with the reference's interface (semilearn/core/utils/misc.py:131-164: register / load /
    update / apply_shadow / restore, `shadow` dict by parameter name).  The shadow tensors ARE the parameters of
    `ema_model`, so after update() the EMA model is current without the reference's two load_state_dict passes per step
    (core/hooks/ema.py:23-24); update() is one srw_ema_step launch over all tensors."""

    def __init__(self, model, decay, ema_model=None):
        self.model, self.decay, self.ema_model = model, float(decay), ema_model
        self.shadow, self.backup, self._table = {}, {}, None

    def _named(self):
        m = self.model.module if hasattr(self.model, "module") else self.model
        return list(m.named_parameters())

    def register(self, keep_ema_values=False):
        """keep_ema_values: the EMA model already holds the shadow (a resumed run: ema.py:17-18 `ema.load(ema_model)`)."""
        named = self._named()
        if self.ema_model is not None:
            em = dict(self.ema_model.named_parameters())
            for n, p in named:
                if em[n].device != p.device:
                    em[n].data = em[n].data.to(p.device)
                if not keep_ema_values:
                    em[n].data.copy_(p.data)
                self.shadow[n] = em[n].data
        else:
            for n, p in named:
                self.shadow[n] = p.data.clone()
        self._table = None

    def load(self, ema_model):
        for n, p in ema_model.named_parameters():
            self.shadow[n].copy_(p.data.to(self.shadow[n].device))

    def _build(self):
        named = self._named()
        rows = (L.EmaRow * len(named))()
        blk = 0
        for i, (n, p) in enumerate(named):
            sh = self.shadow[n]
            assert p.is_contiguous() and sh.is_contiguous() and p.dtype == torch.float32
            rows[i].param, rows[i].shadow, rows[i].numel, rows[i].first_block = p.data_ptr(), sh.data_ptr(), p.numel(), blk
            blk += (p.numel() + L.ADAMW_BLOCK_ELEMS - 1) // L.ADAMW_BLOCK_ELEMS
        dev = named[0][1].device
        host = torch.empty(C.sizeof(rows), dtype=torch.uint8)
        C.memmove(host.data_ptr(), C.addressof(rows), C.sizeof(rows))
        self._table = host.to(dev)
        self._n, self._blocks = len(named), blk
        self._key = tuple(p.data_ptr() for _, p in named)

    @torch.no_grad()
    def update(self):
        if self._table is None or self._key != tuple(p.data_ptr() for _, p in self._named()):
            self._build()
        a = L.EmaArgs(num_tensors=self._n, total_blocks=self._blocks, table=self._table.data_ptr(), decay=self.decay)
        L.check(L.load().srw_ema_step(C.byref(a), L.stream_ptr()), "srw_ema_step")

    def apply_shadow(self):
        for n, p in self._named():
            self.backup[n] = p.data
            p.data = self.shadow[n]

    def restore(self):
        for n, p in self._named():
            p.data = self.backup[n]
        self.backup = {}


class EMAHook(Hook):
    """semilearn/core/hooks/ema.py:8-24: creates algorithm.ema in before_run, updates it after every train step.  The
    shadow lives in algorithm.ema_model's parameters (see EMA)."""

    def before_run(self, algorithm):
        algorithm.ema = EMA(algorithm.model, algorithm.ema_m, ema_model=algorithm.ema_model)
        algorithm.ema.register(keep_ema_values=bool(getattr(algorithm, "resume", False)) and getattr(algorithm, "it", 0) > 0)

    def after_train_step(self, algorithm):
        if getattr(algorithm, "ema", None) is None:
            self.before_run(algorithm)
        algorithm.ema.update()


class PseudoLabelingHook(Hook):
    """gen_ulb_targets: hard labels = argmax; soft labels = softmax(logits / T).  On the fused path the hard labels come
    out of srw_flexmatch_mask (`algorithm._last_pseudo`) so no extra kernel runs; the generic branch stays available."""

    @torch.no_grad()
    def gen_ulb_targets(self, algorithm, logits, use_hard_label=True, T=1.0, softmax=True, label_smoothing=0.0):
        if use_hard_label:
            cached = getattr(algorithm, "_last_pseudo", None)
            if cached is not None and cached[0].data_ptr() == logits.data_ptr():
                return cached[1]
            raise RuntimeError("hard pseudo-labels are produced by the fused masking kernel; call the MaskingHook first")
        raise NotImplementedError("soft pseudo-labels (hard_label: False) are not used by any SemiReward config")


class FlexMatchThresholdingHook(Hook):
    """Device-resident FlexMatch state: selected_label int64[ulb_dest_len] (-1 = never selected), classwise_acc fp32[C],
    plus hist int32[C+1] replacing the reference's host Counter (utils.py:25-29)."""

    def __init__(self, ulb_dest_len, num_classes, thresh_warmup=True, device="cuda"):
        self.ulb_dest_len, self.num_classes, self.thresh_warmup = int(ulb_dest_len), int(num_classes), bool(thresh_warmup)
        self.device = torch.device(device)
        self.selected_label = torch.full((self.ulb_dest_len,), -1, dtype=torch.long, device=self.device)
        self.classwise_acc = torch.zeros(self.num_classes, dtype=torch.float32, device=self.device)
        self._hist = None
        self._rebuild_hist()

    def _rebuild_hist(self):
        """Recount after selected_label was replaced from outside (checkpoint load)."""
        h = torch.bincount(self.selected_label + 1, minlength=self.num_classes + 1).to(torch.int32)
        self._hist = h.contiguous()
        self._hist_src = self.selected_label.data_ptr()

    @torch.no_grad()
    def masking(self, algorithm, logits_x_ulb, idx_ulb, softmax_x_ulb=True, raw_logits=None, *args, **kwargs):
        """logits_x_ulb: raw logits (softmax_x_ulb=True) — on the fused path the algorithm passes the raw weak logits and
        receives (mask, probs, pseudo) computed in one launch."""
        if not softmax_x_ulb:
            raise RuntimeError("fused FlexMatch hook takes raw logits (softmax is fused into the kernel)")
        if self.selected_label.device != logits_x_ulb.device:
            self.selected_label = self.selected_label.to(logits_x_ulb.device)
            self.classwise_acc = self.classwise_acc.to(logits_x_ulb.device)
            self._rebuild_hist()
        if self._hist_src != self.selected_label.data_ptr():
            self._rebuild_hist()
        lw = logits_x_ulb.detach()
        if lw.stride(-1) != 1:
            lw = lw.contiguous()
        B, Cn = lw.shape
        dev = lw.device
        probs = torch.empty(B, Cn, dtype=torch.float32, device=dev)
        pseudo = torch.empty(B, dtype=torch.long, device=dev)
        mask = torch.empty(B, dtype=torch.float32, device=dev)
        idx = idx_ulb.to(device=dev, dtype=torch.long).contiguous()
        a = L.FlexMatchMaskArgs(B=B, num_classes=Cn, ulb_dest_len=self.ulb_dest_len, logits_w=lw.data_ptr(), ld_logits=lw.stride(0),
                                idx_ulb=idx.data_ptr(), p_cutoff=float(algorithm.p_cutoff), thresh_warmup=int(self.thresh_warmup),
                                selected_label=self.selected_label.data_ptr(), hist=self._hist.data_ptr(),
                                classwise_acc=self.classwise_acc.data_ptr(), probs_w=probs.data_ptr(), pseudo=pseudo.data_ptr(),
                                mask=mask.data_ptr(), max_probs=None)
        L.check(L.load().srw_flexmatch_mask(C.byref(a), L.stream_ptr()), "srw_flexmatch_mask")
        algorithm._last_pseudo = (probs, pseudo)
        algorithm._last_probs = probs
        return mask


class FixedThresholdingHook(Hook):
    """semilearn/algorithms/hooks/masking.py:42-57 (FixMatch / UDA / pseudo-label): mask = max_p >= p_cutoff, stateless.
    Softmax + hard pseudo-labels + mask in one launch (srw_flexmatch_mask without FlexMatch state)."""

    @torch.no_grad()
    def masking(self, algorithm, logits_x_ulb, softmax_x_ulb=True, *args, **kwargs):
        if not softmax_x_ulb:
            raise RuntimeError("fused FixedThresholdingHook takes raw logits (softmax is fused into the kernel)")
        lw = _contig_logits(logits_x_ulb)
        B, Cn = lw.shape
        dev = lw.device
        probs = torch.empty(B, Cn, dtype=torch.float32, device=dev)
        pseudo = torch.empty(B, dtype=torch.long, device=dev)
        mask = torch.empty(B, dtype=torch.float32, device=dev)
        a = L.FlexMatchMaskArgs(B=B, num_classes=Cn, ulb_dest_len=0, logits_w=lw.data_ptr(), ld_logits=lw.stride(0), idx_ulb=None,
                                p_cutoff=float(algorithm.p_cutoff), thresh_warmup=0, selected_label=None, hist=None, classwise_acc=None,
                                probs_w=probs.data_ptr(), pseudo=pseudo.data_ptr(), mask=mask.data_ptr(), max_probs=None)
        L.check(L.load().srw_flexmatch_mask(C.byref(a), L.stream_ptr()), "srw_flexmatch_mask")
        algorithm._last_pseudo = (probs, pseudo)
        algorithm._last_probs = probs
        return mask


def _contig_logits(logits_x_ulb):
    lw = logits_x_ulb.detach()
    return lw if lw.stride(-1) == 1 else lw.contiguous()


class FreeMatchThresholdingHook(Hook):
    """Device-resident FreeMatch self-adaptive threshold state (semilearn/algorithms/freematch/utils.py:10-66): time_p,
    p_model[C], label_hist[C].  masking() = softmax + update() + mask + hard pseudo-labels in ONE launch
    (srw_freematch_mask); the attributes keep the reference's names, so get_save_dict / load_model work unchanged."""

    def __init__(self, num_classes, momentum=0.999, device="cuda"):
        self.num_classes, self.m = int(num_classes), float(momentum)
        dev = torch.device(device)
        self.p_model = torch.ones(self.num_classes, dtype=torch.float32, device=dev) / self.num_classes
        self.label_hist = torch.ones(self.num_classes, dtype=torch.float32, device=dev) / self.num_classes
        self.time_p = self.p_model.mean().reshape(1)

    @staticmethod
    def _distributed(algorithm):
        return bool(getattr(algorithm, "distributed", False)) and getattr(algorithm, "world_size", 1) > 1

    @torch.no_grad()
    def masking(self, algorithm, logits_x_ulb, softmax_x_ulb=True, pseudo_from_probs=False, probs_all=None, *args, **kwargs):
        """probs_all (testing hook): phase-2 input as if gathered from all ranks; the local probabilities must be its rows
        [rank * B, (rank + 1) * B)."""
        if not softmax_x_ulb:
            raise RuntimeError("fused FreeMatch hook takes raw logits (softmax is fused into the kernel)")
        lw = _contig_logits(logits_x_ulb)
        dev = lw.device
        for n in ("p_model", "label_hist", "time_p"):
            t = getattr(self, n)
            if t.device != dev or not t.is_contiguous() or t.dtype != torch.float32:
                setattr(self, n, t.to(device=dev, dtype=torch.float32).contiguous())
        self.time_p = self.time_p.reshape(1)
        B, Cn = lw.shape
        probs = torch.empty(B, Cn, dtype=torch.float32, device=dev)
        pseudo = torch.empty(B, dtype=torch.long, device=dev)
        mask = torch.empty(B, dtype=torch.float32, device=dev)
        a = L.FreeMatchMaskArgs(B=B, num_classes=Cn, logits_w=lw.data_ptr(), ld_logits=lw.stride(0), momentum=self.m,
                                use_quantile=int(bool(algorithm.use_quantile)), clip_thresh=int(bool(algorithm.clip_thresh)),
                                time_p=self.time_p.data_ptr(), p_model=self.p_model.data_ptr(), label_hist=self.label_hist.data_ptr(),
                                probs_w=probs.data_ptr(), pseudo=pseudo.data_ptr(), pseudo_from_probs=int(bool(pseudo_from_probs)),
                                mask=mask.data_ptr(), max_probs=None, phase=0, probs_all=None, B_all=0)
        if probs_all is not None or self._distributed(algorithm):
            # data parallel: update() sees every rank's probabilities (concat_all_gather, utils.py:25-26; rank-major rows as in
            # semilearn/algorithms/utils/ops.py:34-45): softmax launch, all_gather, then the update + mask launch
            a.phase = 1
            L.check(L.load().srw_freematch_mask(C.byref(a), L.stream_ptr()), "srw_freematch_mask")
            if probs_all is None:
                import torch.distributed as dist
                probs_all = torch.empty(dist.get_world_size() * B, Cn, dtype=torch.float32, device=dev)
                dist.all_gather_into_tensor(probs_all, probs)
            a.phase, a.probs_all, a.B_all = 2, probs_all.data_ptr(), probs_all.shape[0]
        L.check(L.load().srw_freematch_mask(C.byref(a), L.stream_ptr()), "srw_freematch_mask")
        algorithm.p_model, algorithm.label_hist, algorithm.time_p = self.p_model, self.label_hist, self.time_p   # utils.py:41-43
        algorithm._last_pseudo = (probs, pseudo)
        algorithm._last_probs = probs
        return mask


class DistAlignEMAHook(Hook):
    """State holder of DistAlignEMAHook with a uniform target (semilearn/algorithms/hooks/dist_align.py:10-72); the
    alignment itself runs inside srw_softmatch_mask."""

    def __init__(self, num_classes, momentum=0.999, p_target_type="uniform", p_target=None, device="cuda"):
        if p_target_type != "uniform":
            raise NotImplementedError("DistAlign p_target_type 'model'/'gt': every SemiReward SoftMatch config uses dist_uniform: True")
        self.num_classes, self.m = int(num_classes), float(momentum)
        dev = torch.device(device)
        self.p_target = torch.ones(self.num_classes, dtype=torch.float32, device=dev) / self.num_classes
        self._p_model = torch.zeros(self.num_classes, dtype=torch.float32, device=dev)
        self._initialized = torch.zeros(1, dtype=torch.int32, device=dev)

    @property
    def p_model(self):   # None until the first dist_align, like the reference (dist_align.py:22)
        return self._p_model

    @p_model.setter
    def p_model(self, value):   # load_model restores it (srsoftmatch.py:236)
        self._p_model = value.to(device=self._p_model.device, dtype=torch.float32).contiguous()
        self._initialized.fill_(1)


class SoftMatchWeightingHook(Hook):
    """Device-resident SoftMatch truncated-Gaussian weighting state (semilearn/algorithms/srsoftmatch/utils.py:12-77):
    prob_max_mu_t, prob_max_var_t.  masking() = softmax (+ DistAlign) + update() + weights + hard pseudo-labels in ONE
    launch (srw_softmatch_mask), without the reference's two .item() syncs."""

    def __init__(self, num_classes, n_sigma=2, momentum=0.999, per_class=False, device="cuda"):
        if per_class:
            raise NotImplementedError("SoftMatch per_class: True is not used by any SemiReward config")
        self.num_classes, self.n_sigma, self.m = int(num_classes), int(n_sigma), float(momentum)
        dev = torch.device(device)
        self.prob_max_mu_t = torch.full((1,), 1.0 / self.num_classes, dtype=torch.float32, device=dev)
        self.prob_max_var_t = torch.ones(1, dtype=torch.float32, device=dev)

    @torch.no_grad()
    def masking(self, algorithm, logits_x_ulb, softmax_x_ulb=True, dist_align=False, pseudo_from_probs=False, gathered=None,
                *args, **kwargs):
        """gathered (testing hook): callable(kind, local_tensor) -> tensor standing in for the all_gather over ranks
        (kind 'probs' -> [W*B, C], 'max_probs' -> [W*B])."""
        if not softmax_x_ulb:
            raise RuntimeError("fused SoftMatch hook takes raw logits (softmax and DistAlign are fused into the kernel)")
        lw = _contig_logits(logits_x_ulb)
        dev = lw.device
        for n in ("prob_max_mu_t", "prob_max_var_t"):
            t = getattr(self, n)
            if t.device != dev or t.dim() != 1 or t.dtype != torch.float32:
                setattr(self, n, t.to(device=dev, dtype=torch.float32).reshape(1).contiguous())
        da = algorithm.hooks_dict["DistAlignHook"] if dist_align else None
        B, Cn = lw.shape
        probs = torch.empty(B, Cn, dtype=torch.float32, device=dev)
        pseudo = torch.empty(B, dtype=torch.long, device=dev)
        mask = torch.empty(B, dtype=torch.float32, device=dev)
        a = L.SoftMatchMaskArgs(B=B, num_classes=Cn, logits_w=lw.data_ptr(), ld_logits=lw.stride(0), momentum=self.m, n_sigma=self.n_sigma,
                                dist_align=int(da is not None), da_p_model=L.ptr(da._p_model) if da else None,
                                da_p_target=L.ptr(da.p_target) if da else None, da_initialized=L.ptr(da._initialized) if da else None,
                                prob_max_mu_t=self.prob_max_mu_t.data_ptr(), prob_max_var_t=self.prob_max_var_t.data_ptr(),
                                probs_w=probs.data_ptr(), probs_aligned=None, pseudo=pseudo.data_ptr(),
                                pseudo_from_probs=int(bool(pseudo_from_probs)), mask=mask.data_ptr(), max_probs=None,
                                phase=0, probs_all=None, B_all=0, maxp_all=None, n_all=0)
        launch = lambda: L.check(L.load().srw_softmatch_mask(C.byref(a), L.stream_ptr()), "srw_softmatch_mask")  # noqa: E731
        distributed = bool(getattr(algorithm, "distributed", False)) and getattr(algorithm, "world_size", 1) > 1
        if gathered is not None or distributed:
            # data parallel (C4): DistAlign and the weighting statistics see every rank's rows (concat_all_gather,
            # dist_align.py:40-42, srsoftmatch/utils.py:33-34); masks are formed for the local rows
            import torch.distributed as dist

            def gather(kind, t):
                if gathered is not None:
                    return gathered(kind, t)
                out = torch.empty((dist.get_world_size() * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
                dist.all_gather_into_tensor(out, t)
                return out
            maxp = torch.empty(B, dtype=torch.float32, device=dev)
            a.max_probs, a.phase = maxp.data_ptr(), 1
            launch()
            if da is not None:
                probs_all = gather("probs", probs)
                a.phase, a.probs_all, a.B_all = 2, probs_all.data_ptr(), probs_all.shape[0]
                launch()
            maxp_all = gather("max_probs", maxp)
            a.phase, a.maxp_all, a.n_all = 3, maxp_all.data_ptr(), maxp_all.shape[0]
        launch()
        algorithm._last_pseudo = (probs, pseudo)
        algorithm._last_probs = probs
        return mask
