"""SRPseudoLabel — Pseudo-Label + SemiReward train step on the B200-native kernels, registered under the reference's name.

Follows semilearn/algorithms/srpseudolabel/srpseudolabel.py (task_type 'cls'): ctor :35-50, init :51-53, set_hooks :55-58,
data_generator :60-90, train_step :92-201, get_argument :203-218.  One unlabelled view: the consistency loss is taken on
the same logits that produce the pseudo-labels, so labelled AND unlabelled rows carry gradient; the unsupervised term is
ramped by clip(it / (unsup_warm_up * num_train_iter), 0, 1); masks come from the stateless FixedThresholdingHook.
The reference calls the model separately for x_lb and x_ulb_w (BatchNorm nets freeze BN for the second call); a LayerNorm
backbone computes the same numbers in one batched call, which is what runs here.

Stage 2 (it > start_timing): the reference re-runs the model on x_ulb_w K = sr_decay() times and keeps the last pass.  The
hook is stateless, so with DropPath off the K passes are identical and nothing is re-run; with DropPath on only pass 0
(labelled rows -> sup loss; unlabelled rows -> util_ratio, SR update) and pass K (unlabelled rows -> unsup loss) are ever
used, so the step runs [x_lb | x_ulb_w (pass K) | x_ulb_w (pass 0)] as one forward with per-row DropPath draws."""
from __future__ import annotations

import numpy as np
import torch

from ..core.hooks import FixedThresholdingHook, PseudoLabelingHook
from ..core.registry import ALGORITHMS
from .srflexmatch import SRFlexMatch, _PrecomputedGrads, _ssl_loss_native
from .utils import SSL_Argument, str2bool


@ALGORITHMS.register("srpseudolabel")
class SRPseudoLabel(SRFlexMatch):
    def _init_algorithm(self, args):
        if getattr(args, "task_type", "cls") != "cls":
            raise NotImplementedError("SRPseudoLabel task_type 'reg' (L1 consistency on noisy inputs) is not used by any classification config")
        self.init(p_cutoff=args.p_cutoff, unsup_warm_up=args.unsup_warm_up)
        self.task_type = "cls"

    def init(self, p_cutoff, unsup_warm_up=0.4):
        self.p_cutoff, self.unsup_warm_up = p_cutoff, unsup_warm_up

    def set_hooks(self):
        self.register_hook(PseudoLabelingHook(), "PseudoLabelingHook")
        self.register_hook(FixedThresholdingHook(), "MaskingHook")
        super(SRFlexMatch, self).set_hooks()

    def train_step(self, x_lb, y_lb, x_ulb_w):
        net = self._net()
        if not (torch.is_grad_enabled() and hasattr(net, "forward_native")):
            raise RuntimeError("SRPseudoLabel.train_step runs the native eager-backward step only (grad mode on, semireward_b200 ViT)")
        self._sr_wait()
        dev = x_lb.device
        nl, nu = x_lb.shape[0], x_ulb_w.shape[0]
        stage2 = self.it > self.start_timing
        two_pass = stage2 and self._stochastic_backbone()      # pass K's unlabelled rows differ from pass 0's only under DropPath
        parts = [x_lb, x_ulb_w] + ([x_ulb_w] if two_pass else [])
        xb = net.input_buffer((nl + nu * (2 if two_pass else 1),) + tuple(x_lb.shape[1:]), dev)
        torch.cat(parts, out=xb)
        lg, ft, h = net.forward_native(xb, grad_batch=nl + nu)
        logits_lb, feats_lb = lg[:nl], ft[:nl]
        logits_g, feats_g = lg[nl:nl + nu], ft[nl:nl + nu]                  # rows whose logits carry the unsup gradient
        logits_0, feats_0 = (lg[nl + nu:], ft[nl + nu:]) if two_pass else (logits_g, feats_g)   # pass 0's unlabelled rows
        feat_dict = {"x_lb": feats_lb, "x_ulb_w": feats_0}
        y_lb = y_lb.to(torch.long)
        mask = self.call_hook("masking", "MaskingHook", logits_x_ulb=logits_0, softmax_x_ulb=True)
        pseudo_label = self._last_pseudo[1]
        dl = net.dlogits_buffer(nl + nu, dev)
        warm = float(np.clip(self.it / (self.unsup_warm_up * self.num_train_iter), a_min=0.0, a_max=1.0))   # srpseudolabel.py:194
        if stage2:
            if two_pass:
                mask_dg = self.call_hook("masking", "MaskingHook", logits_x_ulb=logits_g, softmax_x_ulb=True)
                pseudo_dg = self._last_pseudo[1]
            else:
                mask_dg, pseudo_dg = mask, pseudo_label
            reward_dg = self.rewarder(feats_g, pseudo_dg)
            losses, mask2 = _ssl_loss_native(logits_lb, logits_g, y_lb, pseudo_dg, mask_dg, reward_dg.view(-1), self.lambda_u * warm, dl[:nl], dl[nl:])
        else:
            losses, mask2 = _ssl_loss_native(logits_lb, logits_g, y_lb, pseudo_label, mask, None, self.lambda_u * warm, dl[:nl], dl[nl:])
        host = self._host_scalars(5 + nu)
        host[:5].copy_(losses, non_blocking=True)
        host[5:].copy_(mask, non_blocking=True)
        copied = torch.cuda.Event()
        copied.record()
        if self.it > 0:
            if self.it >= self.start_timing:
                if self.it % self.N_k == 0 and self.it > self.start_timing:
                    self.max_reward = -float("inf")
                    self._sr_update_async(feats_0, pseudo_label)
            else:
                self._sr_update_async(feats_lb, y_lb)
        net.backward_native(h, dl)
        net.allreduce_grads_()
        total_loss = _PrecomputedGrads.apply(losses[2], net, net.cls_token)
        copied.synchronize()
        sup, unsup, total, _ = host[:4].tolist()
        out_dict = self.process_out_dict(loss=total_loss, feat=feat_dict)
        log_dict = self.process_log_dict(sup_loss=sup, unsup_loss=unsup, total_loss=total, util_ratio=float(host[5:].mean()))
        self._last_mask, self._last_mask2, self._last_pseudo_label = mask, mask2, pseudo_label
        return out_dict, log_dict

    def get_save_dict(self):
        d = super(SRFlexMatch, self).get_save_dict()
        d["semireward"] = self._sr_save_dict()
        return d

    def load_model(self, load_path):
        ck = super(SRFlexMatch, self).load_model(load_path)
        self._sr_load(ck)
        return ck

    @staticmethod
    def get_argument():
        return [SSL_Argument("--p_cutoff", float, 0.95), SSL_Argument("--unsup_warm_up", float, 0.4, "warm up ratio for unsupervised loss"),
                SSL_Argument("--task_type", str, "cls"), SSL_Argument("--start_timing", int, 20000), SSL_Argument("--feature_dim", int, 384),
                SSL_Argument("--sr_lr", float, 0.0005), SSL_Argument("--N_k", int, 10), SSL_Argument("--sr_ema", str2bool, True),
                SSL_Argument("--sr_ema_m", float, 0.999), SSL_Argument("--range", int, 100)]
