"""SRFlexMatch — FlexMatch + SemiReward train step on the B200-native kernels, registered under the reference's name.

Follows semilearn/algorithms/srflexmatch/srflexmatch.py: ctor :44-58, set_hooks :66-70, data_generator :72-104,
train_step :107-217, get_save_dict/load_model :219-231, get_argument :233-246.

What runs where (all device work is in libsrw_b200.so; this file only sequences calls):
  backbone fwd/bwd ........ srw_vit_forward / srw_vit_backward  (one autograd.Function; weak rows carry no gradient)
  probs/pseudo/mask/hook .. srw_flexmatch_mask   (1 launch instead of ~10 + a 400 KB D2H Counter round trip)
  reward / mask2 .......... srw_rewarder_fwd + srw_ssl_loss
  sup/unsup/total + dlogits srw_ssl_loss          (loss tensor is returned with a grad_fn so ParamUpdateHook's
                                                   loss.backward() works unchanged)
  SR online update ........ srw_generator_fwd + srw_rewarder_train (fwd + 2 MSE + bwd + Adam in one launch)
  4x .item() .............. one 16-byte D2H copy of the loss vector
The engine wants the gradient-carrying rows first, so the concatenation order is (x_lb, x_ulb_s, x_ulb_w) instead of
the reference's (x_lb, x_ulb_w, x_ulb_s); per-row results are identical because LayerNorm nets do not couple rows."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib as L
from ..core.algorithmbase import AlgorithmBase
from ..core.hooks import FlexMatchThresholdingHook, PseudoLabelingHook
from ..core.registry import ALGORITHMS
from .semireward import EMARewarder, Generator, Rewarder, label_dim
from .utils import SSL_Argument, str2bool


def _nrows(x):
    """Batch rows of an input: image tensor [B, ...] or a text dict {'input_ids': [B, L], ...}."""
    return (x["input_ids"] if isinstance(x, dict) else x).shape[0]


def _device_of(x):
    return (x["input_ids"] if isinstance(x, dict) else x).device


class _SSLLoss(torch.autograd.Function):
    """total_loss with a grad_fn: forward evaluates srw_ssl_loss (which also writes d total / d logits), backward hands
    those gradients to the two logit tensors (they may come from two different backbone passes in stage 2)."""

    @staticmethod
    def forward(ctx, logits_lb, logits_s, y_lb, pseudo, mask, reward, lambda_u, out):
        B_lb, Cn = logits_lb.shape
        B_u = logits_s.shape[0]
        dev = logits_lb.device
        dl_lb = torch.empty(B_lb, Cn, dtype=torch.float32, device=dev)
        dl_s = torch.empty(B_u, Cn, dtype=torch.float32, device=dev)
        losses = torch.empty(4, dtype=torch.float32, device=dev)
        mask2 = torch.empty(B_u, dtype=torch.float32, device=dev)
        assert logits_lb.stride(1) == 1 and logits_s.stride(1) == 1 and logits_lb.stride(0) == logits_s.stride(0)
        a = L.SslLossArgs(B_lb=B_lb, B_ulb=B_u, num_classes=Cn, logits_lb=logits_lb.data_ptr(), logits_s=logits_s.data_ptr(),
                          ld_logits=logits_lb.stride(0), y_lb=y_lb.data_ptr(), pseudo=pseudo.data_ptr(), mask=mask.data_ptr(),
                          reward=L.ptr(reward), lambda_u=float(lambda_u), mask2=mask2.data_ptr(), losses=losses.data_ptr(),
                          dlogits_lb=dl_lb.data_ptr(), dlogits_s=dl_s.data_ptr(), ld_dlogits=Cn)
        L.check(L.load().srw_ssl_loss(C.byref(a), L.stream_ptr()), "srw_ssl_loss")
        ctx.dl_lb, ctx.dl_s = dl_lb, dl_s
        out["losses"], out["mask2"] = losses, mask2
        return losses[2].clone()

    @staticmethod
    def backward(ctx, g):
        return ctx.dl_lb * g, ctx.dl_s * g, None, None, None, None, None, None


def _ssl_loss_native(logits_lb, logits_s, y_lb, pseudo, mask, reward, lambda_u, dl_lb, dl_s, rewarder=None, feats=None):
    """srw_ssl_loss -> (losses[4] = sup, unsup, total, util ; mask2).  d total / d logits goes into dl_lb / dl_s (row stride C).
    rewarder + feats: the fused stage-2 epilogue — Rewarder.forward(feats, pseudo), the mean-threshold mask2, the masked
    consistency loss and d loss / d logits in ONE launch (`reward` is ignored then)."""
    B_lb, Cn = logits_lb.shape
    B_u = logits_s.shape[0]
    dev = logits_lb.device
    losses = torch.empty(5, dtype=torch.float32, device=dev)   # [4]: extra term of the algorithm (FreeMatch entropy)
    mask2 = torch.empty(B_u, dtype=torch.float32, device=dev)
    assert logits_lb.stride(1) == 1 and logits_s.stride(1) == 1 and logits_lb.stride(0) == logits_s.stride(0)
    a = L.SslLossArgs(B_lb=B_lb, B_ulb=B_u, num_classes=Cn, logits_lb=logits_lb.data_ptr(), logits_s=logits_s.data_ptr(),
                      ld_logits=logits_lb.stride(0), y_lb=y_lb.data_ptr(), pseudo=pseudo.data_ptr(), mask=mask.data_ptr(),
                      reward=L.ptr(reward), lambda_u=float(lambda_u), mask2=mask2.data_ptr(), losses=losses.data_ptr(),
                      dlogits_lb=dl_lb.data_ptr(), dlogits_s=dl_s.data_ptr(), ld_dlogits=Cn)
    keep = None
    if rewarder is not None:
        f = feats.detach()
        if f.stride(-1) != 1:
            f = f.contiguous()
        ws = torch.empty(max(L.load().srw_rewarder_workspace_floats(B_u, rewarder.feature_dim), 16), dtype=torch.float32, device=dev)
        keep = (f, ws, L.ptr_array(rewarder._params()))
        a.rp, a.feats, a.ld_feats, a.feature_dim, a.label_rows, a.rew_workspace = keep[2], f.data_ptr(), f.stride(0), rewarder.feature_dim, rewarder.label_rows, ws.data_ptr()
        a.reward = None
    L.check(L.load().srw_ssl_loss(C.byref(a), L.stream_ptr()), "srw_ssl_loss")
    if rewarder is not None and hasattr(rewarder, "update_ema"):
        rewarder.update_ema()
    return losses, mask2


class _PrecomputedGrads(torch.autograd.Function):
    """The loss tensor handed to ParamUpdateHook when the backbone backward was already launched inside train_step
    (eager backward).  loss.backward() applies the upstream gradient (exactly 1 for a plain `loss.backward()`;
    srw_scale_inplace is a no-op then) and hands the finished gradients to the parameters: `p.grad = view of the flat
    gradient buffer` (or `p.grad += view` when a gradient is already there).  The hand-over is done here instead of through
    152 AccumulateGrad nodes, which would copy every gradient (85.7 MB per step) because the views are long-lived."""

    @staticmethod
    def forward(ctx, loss_value, net, anchor):
        ctx.net = net
        return loss_value.clone()

    @staticmethod
    def backward(ctx, g):
        net = ctx.net
        flat, views = net._flat_grads, net._grad_views
        g = g.to(torch.float32).contiguous()
        L.check(L.load().srw_scale_inplace(flat.data_ptr(), flat.numel(), g.data_ptr(), L.stream_ptr()), "srw_scale_inplace")
        for p, v in zip(net._grad_params(), views):
            if p.grad is None:
                p.grad = v
            else:
                p.grad.add_(v)
        return None, None, None


@ALGORITHMS.register("srflexmatch")
class SRFlexMatch(AlgorithmBase):
    def __init__(self, args, net_builder, tb_log=None, logger=None):
        super().__init__(args, net_builder, tb_log, logger)
        self._init_algorithm(args)
        self.N_k = args.N_k
        if args.sr_ema != 0:   # srflexmatch.py:49-50
            self.rewarder = EMARewarder(label_dim(self.num_classes), 128, feature_dim=args.feature_dim, ema_decay=args.sr_ema_m)
        else:
            self.rewarder = Rewarder(label_dim(self.num_classes), 128, args.feature_dim)
        self.generator = Generator(args.feature_dim)
        if torch.cuda.is_available():   # reference: send_model_cuda(args, Rewarder(...)) in the ctor (srflexmatch.py:49-51)
            self.rewarder, self.generator = self.rewarder.cuda(self.gpu), self.generator.cuda(self.gpu)
            if self.distributed:
                import torch.distributed as dist
                with torch.no_grad():
                    for p in list(self.rewarder.parameters()) + list(self.generator.parameters()):
                        dist.broadcast(p.data, src=0)
                self.rewarder._dp_group = dist.group.WORLD
        self.start_timing = args.start_timing
        self.sr_lr = args.sr_lr
        self.max_reward = -float("inf")
        # with DropPath/dropout off every sampling pass of data_generator sees identical logits: run the backbone once
        # and replay only the hook-state updates (bit-equivalent in that mode only — SURVEY.md §8a a2)
        self.replay_deterministic_passes = True
        # The SR online update (Generator forward + fused Rewarder train kernel) only consumes detached features and does
        # not feed this step's loss, so it runs on a side stream concurrently with the backbone backward (two single-CTA
        # kernels next to 147 busy SMs).  The next step's first use of the Rewarder waits for it.
        self._sr_stream = None
        self._sr_done = None
        # Eager backward: train_step launches the backbone backward itself as soon as d loss / d logits exists (it is a
        # by-product of the loss kernel), BEFORE the host reads the loss values back for log_dict.  The device therefore
        # never waits for the host between forward and backward; ParamUpdateHook's loss.backward() just collects the
        # gradients.  Set to False to go through autograd (_VitFunction / _SSLLoss) instead.
        self.eager_backward = True
        # Stage 2 with DropPath on: the reference runs 1 + K full backbone passes per step (srflexmatch.py:72-104), but the
        # passes do not depend on each other (only the hook state, which consumes their weak logits in order, does), and of
        # passes 1..K-1 only the weak rows are ever used, of pass K only weak + strong.  batch_stochastic_passes runs all rows
        # that are used (nl + 2 nu + (K + 1) nu instead of (K + 1)(nl + 2 nu)) as ONE forward with per-row DropPath draws and
        # ONE backward over the rows that carry gradient.  Row-for-row identical to the sequential passes given the same
        # DropPath multipliers (tested); False runs the passes one after the other.
        self.batch_stochastic_passes = True

    def _init_algorithm(self, args):
        self.init(T=args.T, p_cutoff=args.p_cutoff, hard_label=args.hard_label, thresh_warmup=args.thresh_warmup)

    def init(self, T, p_cutoff, hard_label=True, thresh_warmup=True):
        self.T, self.p_cutoff, self.use_hard_label, self.thresh_warmup = T, p_cutoff, hard_label, thresh_warmup

    def set_hooks(self):
        self.register_hook(PseudoLabelingHook(), "PseudoLabelingHook")
        self.register_hook(FlexMatchThresholdingHook(ulb_dest_len=self.args.ulb_dest_len, num_classes=self.num_classes,
                                                     thresh_warmup=self.args.thresh_warmup,
                                                     device=f"cuda:{self.gpu}" if torch.cuda.is_available() else "cpu"), "MaskingHook")
        super().set_hooks()

    # -- pieces -----------------------------------------------------------------------------------
    def _backbone(self, x_lb, x_ulb_w, x_ulb_s, need_grad=True):
        """-> logits/feats of (lb, weak, strong).  Gradient rows (lb, strong) go first inside the engine."""
        # use_cat: False makes the reference call the model three times (lb, strong with grad; weak under no_grad,
        # srflexmatch.py:119-130).  For a LayerNorm backbone rows do not interact, so the one batched call below computes the
        # same numbers; the weak rows carry no gradient either way.  (The BERT / HuBERT wrappers that need use_cat: False for
        # their dict / ragged inputs are not built: get_net_builder raises for them.)
        nl, nu = x_lb.shape[0], x_ulb_s.shape[0]
        inputs = torch.cat((x_lb, x_ulb_s, x_ulb_w))
        if need_grad:
            out = self.model(inputs, grad_batch=nl + nu)
        else:
            with torch.no_grad():
                out = self.model(inputs)
        lg, ft = out["logits"], out["feat"]
        return (lg[:nl], lg[nl + nu:], lg[nl:nl + nu]), (ft[:nl], ft[nl + nu:], ft[nl:nl + nu])

    def _stochastic_backbone(self):
        m = self.model.module if hasattr(self.model, "module") else self.model
        if hasattr(m, "stochastic"):
            return m.stochastic()
        return m.training and max(getattr(m, "drop_path_rates", [0.0])) > 0.0

    def _mask_and_pseudo(self, logits_w, idx_ulb, first_pass=True):
        """(mask, hard pseudo-labels) of the weak logits and the MaskingHook state update, one launch.  first_pass: the
        call in train_step itself (True) or one of data_generator's sampling passes (False) — the same thing for FlexMatch."""
        mask = self.call_hook("masking", "MaskingHook", logits_x_ulb=logits_w, idx_ulb=idx_ulb, softmax_x_ulb=True)
        pseudo = self.call_hook("gen_ulb_targets", "PseudoLabelingHook", logits=self._last_probs, use_hard_label=self.use_hard_label,
                                T=self.T, softmax=False)
        return mask, pseudo

    def data_generator(self, x_lb, y_lb, idx_ulb, x_ulb_w, x_ulb_s, rewarder, gpu, first_pass=None):
        """K = sr_decay() sampling passes; the last pass's (logits_s, pseudo, mask, reward) define the loss
        (srflexmatch.py:72-104).  Returns those tensors; the loss itself is formed by _SSLLoss."""
        K = self.sr_decay()
        stochastic = self._stochastic_backbone() or not self.replay_deterministic_passes
        last = None
        for k in range(K):
            if stochastic:
                is_last = k == K - 1
                (l_lb, l_w, l_s), (_, f_w, _) = self._backbone(x_lb, x_ulb_w, x_ulb_s, need_grad=is_last)
            else:
                (l_lb, l_w, l_s), (_, f_w, _) = first_pass
            mask, pseudo = self._mask_and_pseudo(l_w, idx_ulb)
            if stochastic or k == K - 1:   # the reward of pass k only feeds pass k's loss, and only the last loss survives
                last = (l_s, pseudo, mask, rewarder(f_w, pseudo))
        return last

    # -- the step ---------------------------------------------------------------------------------
    def _sr_update_async(self, feats, true_labels):
        main = torch.cuda.current_stream()
        if self._sr_stream is None:
            self._sr_stream = torch.cuda.Stream()
        side = self._sr_stream
        side.wait_stream(main)                       # feats / labels are produced on the main stream
        with torch.cuda.stream(side):
            gen = self.generator.generate_labels(feats)
            self.rewarder.train_step(feats, gen, true_labels, self.sr_lr, self.num_classes)
            for t in (feats, true_labels):
                t.record_stream(side)
        self._sr_done = side.record_event()

    def _sr_wait(self):
        if self._sr_done is not None:
            torch.cuda.current_stream().wait_event(self._sr_done)
            self._sr_done = None

    def _net(self):
        m = self.model
        if hasattr(m, "module"):
            inner = m.module
            # the reference's train.py wraps alg.model with torch DDP (send_model_cuda, misc.py:56-64).  The native step never
            # calls the wrapper's forward, so DDP's reducer hooks never fire: take its process group and average the flat
            # gradient buffer ourselves (same all-reduce(avg), one message)
            if getattr(inner, "_dp_group", None) is None and hasattr(m, "process_group"):
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and dist.get_world_size(m.process_group) > 1:
                    inner._dp_group = m.process_group
            return inner
        return m

    def _host_scalars(self, n):
        """Two alternating pinned staging buffers for the per-step D2H read of the loss vector."""
        bufs = getattr(self, "_host_bufs", None)
        if bufs is None or bufs[0].numel() != n:
            bufs = self._host_bufs = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._host_flip = 0
        self._host_flip ^= 1
        return bufs[self._host_flip]

    def _backbone_native(self, x_lb, x_ulb_w, x_ulb_s, need_grad=True, drop_scale=None):
        """Autograd-free pass on the net's persistent buffers -> (logits, feats, handle), split as (lb, weak, strong).
        drop_scale: the launch's stochastic-regularisation streams in engine row order (lb, strong, weak) from net.streams_for()
        (DropPath multipliers for a ViT, dropout stream keys for BERT), or None to draw one fresh pass."""
        # use_cat: False makes the reference call the model three times (lb, strong with grad; weak under no_grad,
        # srflexmatch.py:119-130).  For a LayerNorm backbone rows do not interact, so the one batched launch below computes the
        # same numbers; the weak rows carry no gradient either way.  Text batches are dicts {'input_ids', 'attention_mask'}.
        net = self._net()
        nl, nu = _nrows(x_lb), _nrows(x_ulb_s)
        dev = _device_of(x_lb)
        if drop_scale is None and net.stochastic():
            drop_scale = net.streams_for(net.draw_streams(1, nl, nu, dev), [(0, "lb"), (0, "s"), (0, "w")], nl, nu, dev)
        xb = net.concat_inputs([x_lb, x_ulb_s, x_ulb_w], dev)
        lg, ft, handle = net.forward_native(xb, grad_batch=(nl + nu) if need_grad else 0, drop_scale=drop_scale)
        return (lg[:nl], lg[nl + nu:], lg[nl:nl + nu]), (ft[:nl], ft[nl + nu:], ft[nl:nl + nu]), handle

    def _has_extra_loss(self):
        return type(self)._extra_loss is not SRFlexMatch._extra_loss

    def _train_step_stage2_batched(self, x_lb, y_lb, idx_ulb, x_ulb_w, x_ulb_s):
        """Stage 2 with a stochastic backbone: every row the step uses, in one forward and one backward (see
        batch_stochastic_passes in the ctor).  Engine rows: [lb (pass 0) | strong (pass K) | strong (pass 0) | weak (pass 0..K)];
        the first two groups (three when the algorithm adds a loss term on pass 0's strong logits) carry gradient."""
        self._sr_wait()
        net, dev = self._net(), _device_of(x_lb)
        nl, nu, K = _nrows(x_lb), _nrows(x_ulb_s), self.sr_decay()
        # the streams (DropPath multipliers / dropout keys) of every (pass, row) the reference would draw, passes in order
        draws = net.draw_streams(K + 1, nl, nu, dev)
        ds = net.streams_for(draws, [(0, "lb"), (K, "s"), (0, "s")] + [(k, "w") for k in range(K + 1)], nl, nu, dev)
        xb = net.concat_inputs([x_lb, x_ulb_s, x_ulb_s] + [x_ulb_w] * (K + 1), dev)
        extra = self._has_extra_loss()
        gb = nl + (2 * nu if extra else nu)
        lg, ft, h = net.forward_native(xb, grad_batch=gb, drop_scale=ds)
        w0 = nl + 2 * nu
        logits_lb, l_sK, logits_s0 = lg[:nl], lg[nl:nl + nu], lg[nl + nu:w0]
        l_w = [lg[w0 + k * nu:w0 + (k + 1) * nu] for k in range(K + 1)]
        f_w = [ft[w0 + k * nu:w0 + (k + 1) * nu] for k in range(K + 1)]
        feat_dict = {"x_lb": ft[:nl], "x_ulb_w": f_w[0], "x_ulb_s": ft[nl + nu:w0]}
        y_lb = y_lb.to(torch.long)
        mask, pseudo_label = self._mask_and_pseudo(l_w[0], idx_ulb, first_pass=True)
        for k in range(1, K + 1):
            mask_dg, pseudo_dg = self._mask_and_pseudo(l_w[k], idx_ulb, first_pass=False)
        dl = net.dlogits_buffer(gb, dev)
        # fused stage-2 epilogue: Rewarder(f_w of the last pass, its pseudo-labels) -> mask2 -> masked consistency loss -> dlogits, one launch
        losses, mask2 = _ssl_loss_native(logits_lb, l_sK, y_lb, pseudo_dg, mask_dg, None, self.lambda_u, dl[:nl], dl[nl:nl + nu],
                                         rewarder=self.rewarder, feats=f_w[K])
        if extra:
            dl[nl + nu:].zero_()
            self._extra_loss(losses, mask, logits_s0, dl[nl + nu:])
        host = self._host_scalars(5 + nu)
        host[:5].copy_(losses, non_blocking=True)
        host[5:].copy_(mask, non_blocking=True)
        copied = torch.cuda.Event()
        copied.record()
        if self.it % self.N_k == 0:
            self.max_reward = -float("inf")
            self._sr_update_async(f_w[0], pseudo_label)
        net.backward_native(h, dl)
        net.allreduce_grads_()
        total_loss = _PrecomputedGrads.apply(losses[2], net, net._grad_params()[0])
        copied.synchronize()
        sup, unsup, total, _ = host[:4].tolist()
        out_dict = self.process_out_dict(loss=total_loss, feat=feat_dict)
        log_dict = self.process_log_dict(sup_loss=sup, unsup_loss=unsup, total_loss=total, util_ratio=float(host[5:].mean()))
        self._last_mask, self._last_mask2, self._last_pseudo_label = mask, mask2, pseudo_label
        self._last_mask_dg, self._last_pseudo_dg = mask_dg, pseudo_dg   # the last sampling pass (what the unsupervised loss used)
        return out_dict, log_dict

    def _train_step_eager(self, x_lb, y_lb, idx_ulb, x_ulb_w, x_ulb_s):
        """Same step as train_step's autograd route (srflexmatch.py:107-217), with the backward launched in here."""
        stochastic2 = self.it > self.start_timing and (self._stochastic_backbone() or not self.replay_deterministic_passes)
        if stochastic2 and self.batch_stochastic_passes:
            return self._train_step_stage2_batched(x_lb, y_lb, idx_ulb, x_ulb_w, x_ulb_s)
        self._sr_wait()
        net = self._net()
        dev = _device_of(x_lb)
        nl, nu = _nrows(x_lb), _nrows(x_ulb_s)
        full = None
        one_pass = [(0, "lb"), (0, "s"), (0, "w")]
        if stochastic2 and self._stochastic_backbone():   # same draws as the batched route: every pass of the step up front
            full = net.draw_streams(self.sr_decay() + 1, nl, nu, dev)
        if self.it > self.start_timing and not stochastic2 and hasattr(net, "stat_repeats_next"):
            # BatchNorm backbone, deterministic passes: the K sampling passes recompute the identical forward and only advance the
            # running statistics (wrn.py:33,37,97) — the one forward below advances them 1 + K times (srw_wrn_fwd_args.stat_repeats)
            net.stat_repeats_next = self.sr_decay()
        (logits_lb, logits_w, logits_s), (feats_lb, feats_w, feats_s), h0 = self._backbone_native(
            x_lb, x_ulb_w, x_ulb_s, drop_scale=None if full is None else net.streams_for(full, one_pass, nl, nu, dev))
        feat_dict = {"x_lb": feats_lb, "x_ulb_w": feats_w, "x_ulb_s": feats_s}
        y_lb = y_lb.to(torch.long)
        mask, pseudo_label = self._mask_and_pseudo(logits_w, idx_ulb, first_pass=True)
        dl = net.dlogits_buffer(nl + nu, dev)
        dl_lb, dl_s = dl[:nl], dl[nl:]
        h_last = None
        if self.it > self.start_timing:
            K = self.sr_decay()
            stochastic = self._stochastic_backbone() or not self.replay_deterministic_passes
            l_s = logits_s
            for k in range(K):
                if stochastic:   # the reference re-runs the backbone every pass; only the last pass's graph survives
                    (_, l_w, l_s), (_, f_w, _), h = self._backbone_native(
                        x_lb, x_ulb_w, x_ulb_s, need_grad=(k == K - 1),
                        drop_scale=None if full is None else net.streams_for(full, [(k + 1, pt) for _, pt in one_pass], nl, nu, dev))
                    h_last = h if k == K - 1 else None
                else:
                    l_w, f_w = logits_w, feats_w
                mask_dg, pseudo_dg = self._mask_and_pseudo(l_w, idx_ulb, first_pass=False)
            # the reward of pass k only feeds pass k's loss and only the last loss survives: one fused launch for the last pass
            losses, mask2 = _ssl_loss_native(logits_lb, l_s, y_lb, pseudo_dg, mask_dg, None, self.lambda_u, dl_lb, dl_s,
                                             rewarder=self.rewarder, feats=f_w)
        else:
            losses, mask2 = _ssl_loss_native(logits_lb, logits_s, y_lb, pseudo_label, mask, None, self.lambda_u, dl_lb, dl_s)
        # gradient buffers of the pass(es) that carry gradient; algorithm-specific extra terms (FreeMatch's entropy loss on
        # pass 0's strong logits) are added to losses[2] and to the pass-0 gradient here
        if h_last is None:
            self._extra_loss(losses, mask, logits_s, dl_s)
            dl0 = dl1 = None
        else:   # two graphs carry gradient (srflexmatch.py:132 sup through pass 0, :102 unsup through the last pass)
            dl0 = torch.zeros_like(dl)
            dl0[:nl].copy_(dl_lb)
            dl1 = dl.clone()
            dl1[:nl].zero_()
            self._extra_loss(losses, mask, logits_s, dl0[nl:])
        # The loss values (and pass-0's mask for util_ratio) start their way to the host NOW, ahead of the backward in
        # stream order, into pinned memory; the host waits for that copy only after everything else is queued.
        host = self._host_scalars(5 + nu)
        host[:5].copy_(losses, non_blocking=True)
        host[5:].copy_(mask, non_blocking=True)
        copied = torch.cuda.Event()
        copied.record()
        # SR online update first (side stream, forks from here), then the backbone backward on the main stream
        if self.it > 0:
            if self.it >= self.start_timing:
                if self.it % self.N_k == 0 and self.it > self.start_timing:
                    self.max_reward = -float("inf")
                    self._sr_update_async(feats_w, pseudo_label)
            else:
                self._sr_update_async(feats_lb, y_lb)
        if h_last is None:
            net.backward_native(h0, dl)
        else:
            net.backward_native(h0, dl0, final=False)
            net.backward_native(h_last, dl1, accumulate=True)
        net.allreduce_grads_()
        total_loss = _PrecomputedGrads.apply(losses[2], net, net._grad_params()[0])   # one parameter anchors the node in the graph
        copied.synchronize()                        # the device is busy with the backward while the host reads these
        sup, unsup, total, util = host[:4].tolist()
        if self.it > self.start_timing:
            util = float(host[5:].mean())           # the reference logs pass-0's mask (srflexmatch.py:216)
        out_dict = self.process_out_dict(loss=total_loss, feat=feat_dict)
        log_dict = self.process_log_dict(sup_loss=sup, unsup_loss=unsup, total_loss=total, util_ratio=float(util))
        self._last_mask, self._last_mask2, self._last_pseudo_label = mask, mask2, pseudo_label
        self._last_mask_dg, self._last_pseudo_dg = (mask_dg, pseudo_dg) if self.it > self.start_timing else (None, None)
        return out_dict, log_dict

    def _extra_loss(self, losses, mask, logits_s, dl_s):
        """Hook for algorithm-specific loss terms on pass 0's strong logits: add to losses[2] and accumulate into dl_s."""

    def train_step(self, x_lb, y_lb, idx_ulb, x_ulb_w, x_ulb_s):
        if self.eager_backward and torch.is_grad_enabled() and hasattr(self._net(), "forward_native"):
            return self._train_step_eager(x_lb, y_lb, idx_ulb, x_ulb_w, x_ulb_s)
        self._sr_wait()   # last step's Rewarder update must be complete before the Rewarder is read or updated again
        (logits_lb, logits_w, logits_s), (feats_lb, feats_w, feats_s) = self._backbone(x_lb, x_ulb_w, x_ulb_s)
        feat_dict = {"x_lb": feats_lb, "x_ulb_w": feats_w, "x_ulb_s": feats_s}
        y_lb = y_lb.to(torch.long)
        mask, pseudo_label = self._mask_and_pseudo(logits_w, idx_ulb)
        side = {}
        if self.it > self.start_timing:
            l_s, pseudo_dg, mask_dg, reward_dg = self.data_generator(
                x_lb, y_lb, idx_ulb, x_ulb_w, x_ulb_s, self.rewarder, self.gpu,
                first_pass=((logits_lb, logits_w, logits_s), (feats_lb, feats_w, feats_s)))
            total_loss = _SSLLoss.apply(logits_lb, l_s, y_lb, pseudo_dg, mask_dg, reward_dg.view(-1), self.lambda_u, side)
        else:
            total_loss = _SSLLoss.apply(logits_lb, logits_s, y_lb, pseudo_label, mask, None, self.lambda_u, side)
        if self.it > 0:
            if self.it >= self.start_timing:
                # mean reward bookkeeping of srflexmatch.py:165-172 (the "filtered" tensors are always the current batch)
                if self.it % self.N_k == 0 and self.it > self.start_timing:
                    self.max_reward = -float("inf")
                    self._sr_update_async(feats_w.detach(), pseudo_label)
            else:
                self._sr_update_async(feats_lb.detach(), y_lb)
        sup, unsup, total, util = side["losses"].tolist()   # one 16-byte D2H instead of four .item() syncs
        if self.it <= self.start_timing:
            util = float(util)
        else:
            util = float(mask.mean().item())  # the reference logs pass-0's mask (srflexmatch.py:216)
        out_dict = self.process_out_dict(loss=total_loss, feat=feat_dict)
        log_dict = self.process_log_dict(sup_loss=sup, unsup_loss=unsup, total_loss=total, util_ratio=util)
        self._last_mask, self._last_mask2, self._last_pseudo_label = mask, side["mask2"], pseudo_label
        return out_dict, log_dict

    # The reference's checkpoints keep only the backbone and the masking-hook state: a resumed run restarts the Rewarder and the
    # Generator from scratch.  Here the SemiReward state travels under one extra key, "semireward";
    # checkpoints written by the reference (no such key) still load.
    def _sr_save_dict(self):
        self._sr_wait()   # the side-stream online update must have landed before the parameters are read
        return dict(rewarder=self.rewarder.state_dict(), generator=self.generator.state_dict(), rewarder_adam=self.rewarder.optimizer_state())

    def _sr_load(self, ck):
        sr = ck.get("semireward")
        if sr is None:
            return
        self._sr_wait()
        self.rewarder.load_state_dict(sr["rewarder"])
        self.generator.load_state_dict(sr["generator"])
        self.rewarder.load_optimizer_state(sr.get("rewarder_adam"))

    def get_save_dict(self):
        d = super().get_save_dict()
        d["classwise_acc"] = self.hooks_dict["MaskingHook"].classwise_acc.cpu()
        d["selected_label"] = self.hooks_dict["MaskingHook"].selected_label.cpu()
        d["semireward"] = self._sr_save_dict()
        return d

    def load_model(self, load_path):
        ck = super().load_model(load_path)
        h = self.hooks_dict["MaskingHook"]
        h.classwise_acc = ck["classwise_acc"].cuda(self.gpu)
        h.selected_label = ck["selected_label"].cuda(self.gpu)
        h._rebuild_hist()
        self._sr_load(ck)
        return ck

    @staticmethod
    def get_argument():
        return [SSL_Argument("--hard_label", str2bool, True), SSL_Argument("--T", float, 0.5), SSL_Argument("--p_cutoff", float, 0.95),
                SSL_Argument("--thresh_warmup", str2bool, True), SSL_Argument("--start_timing", int, 20000),
                SSL_Argument("--feature_dim", int, 384), SSL_Argument("--sr_lr", float, 0.0005), SSL_Argument("--N_k", int, 10),
                SSL_Argument("--sr_ema", str2bool, True), SSL_Argument("--sr_ema_m", float, 0.999)]
