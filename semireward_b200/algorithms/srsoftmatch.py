"""SRSoftMatch — SoftMatch + SemiReward train step on the B200-native kernels, registered under the reference's name.

Follows semilearn/algorithms/srsoftmatch/srsoftmatch.py: ctor :41-52, init :53-59, data_generator :61-95,
set_hooks :97-106, train_step :108-221, get_save_dict/load_model :223-241, get_argument :243-258.
The step skeleton is SRFlexMatch's; what differs is the hook pair: train_step aligns the weak probabilities with
DistAlignEMAHook before SoftMatchWeightingHook forms its truncated-Gaussian weights (pseudo-labels from the raw
logits), data_generator's sampling passes use the un-aligned probabilities (pseudo-labels from those) — each case is one
srw_softmatch_mask launch with the state on the device (no .item() syncs)."""
from __future__ import annotations

import torch

from ..core.hooks import DistAlignEMAHook, PseudoLabelingHook, SoftMatchWeightingHook
from ..core.registry import ALGORITHMS
from .srflexmatch import SRFlexMatch
from .utils import SSL_Argument, str2bool


@ALGORITHMS.register("srsoftmatch")
class SRSoftMatch(SRFlexMatch):
    def _init_algorithm(self, args):
        self.init(T=args.T, hard_label=args.hard_label, dist_align=args.dist_align, dist_uniform=args.dist_uniform, ema_p=args.ema_p,
                  n_sigma=args.n_sigma, per_class=args.per_class)

    def init(self, T, hard_label=True, dist_align=True, dist_uniform=True, ema_p=0.999, n_sigma=2, per_class=False):
        self.T, self.use_hard_label, self.dist_align, self.dist_uniform = T, hard_label, dist_align, dist_uniform
        self.ema_p, self.n_sigma, self.per_class = ema_p, n_sigma, per_class

    def set_hooks(self):
        dev = f"cuda:{self.gpu}" if torch.cuda.is_available() else "cpu"
        self.register_hook(PseudoLabelingHook(), "PseudoLabelingHook")
        self.register_hook(DistAlignEMAHook(num_classes=self.num_classes, momentum=self.args.ema_p,
                                            p_target_type="uniform" if self.args.dist_uniform else "model", device=dev), "DistAlignHook")
        self.register_hook(SoftMatchWeightingHook(num_classes=self.num_classes, n_sigma=self.args.n_sigma, momentum=self.args.ema_p,
                                                  per_class=self.args.per_class, device=dev), "MaskingHook")
        super(SRFlexMatch, self).set_hooks()

    def _mask_and_pseudo(self, logits_w, idx_ulb, first_pass=True):
        # the reference calls dist_align unconditionally in train_step (srsoftmatch.py:138); data_generator never aligns
        mask = self.call_hook("masking", "MaskingHook", logits_x_ulb=logits_w, softmax_x_ulb=True, dist_align=first_pass,
                              pseudo_from_probs=not first_pass)
        return mask, self._last_pseudo[1]

    def train_step(self, x_lb, y_lb, x_ulb_w, x_ulb_s):
        if not (torch.is_grad_enabled() and hasattr(self._net(), "forward_native")):
            raise RuntimeError("SRSoftMatch.train_step runs the native eager-backward step only (grad mode on, semireward_b200 ViT)")
        return self._train_step_eager(x_lb, y_lb, None, x_ulb_w, x_ulb_s)

    def get_save_dict(self):
        d = super(SRFlexMatch, self).get_save_dict()
        da, h = self.hooks_dict["DistAlignHook"], self.hooks_dict["MaskingHook"]
        d["p_model"], d["p_target"] = da.p_model.cpu(), da.p_target.cpu()
        d["prob_max_mu_t"], d["prob_max_var_t"] = h.prob_max_mu_t.reshape(()).cpu(), h.prob_max_var_t.reshape(()).cpu()   # 0-dim like the reference
        d["semireward"] = self._sr_save_dict()
        return d

    def load_model(self, load_path):
        ck = super(SRFlexMatch, self).load_model(load_path)
        da, h = self.hooks_dict["DistAlignHook"], self.hooks_dict["MaskingHook"]
        da.p_model, da.p_target = ck["p_model"].cuda(self.gpu), ck["p_target"].cuda(self.gpu)
        h.prob_max_mu_t, h.prob_max_var_t = ck["prob_max_mu_t"].cuda(self.gpu), ck["prob_max_var_t"].cuda(self.gpu)
        self._sr_load(ck)
        return ck

    @staticmethod
    def get_argument():
        return [SSL_Argument("--hard_label", str2bool, True), SSL_Argument("--T", float, 0.5), SSL_Argument("--dist_align", str2bool, True),
                SSL_Argument("--dist_uniform", str2bool, True), SSL_Argument("--ema_p", float, 0.999), SSL_Argument("--n_sigma", int, 2),
                SSL_Argument("--per_class", str2bool, False), SSL_Argument("--start_timing", int, 20000),
                SSL_Argument("--feature_dim", int, 384), SSL_Argument("--sr_lr", float, 0.0005), SSL_Argument("--N_k", int, 10),
                SSL_Argument("--sr_ema", str2bool, True), SSL_Argument("--sr_ema_m", float, 0.999)]
