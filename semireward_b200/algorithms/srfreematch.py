"""SRFreeMatch — FreeMatch + SemiReward train step on the B200-native kernels, registered under the reference's name.

Follows semilearn/algorithms/srfreematch/srfreematch.py: ctor :53-70, init :71-76, data_generator :78-111,
set_hooks :113-116, train_step :118-228, entropy_loss :16-44, get_save_dict/load_model :230-245, get_argument :247-262.
The step skeleton (backbone passes, Rewarder, SR online update, eager backward) is SRFlexMatch's; what differs is the
MaskingHook (self-adaptive threshold, one srw_freematch_mask launch), where the hard pseudo-labels come from (logits in
train_step, probabilities in data_generator) and the fairness entropy term srw_freematch_entropy adds to the loss and to
d loss / d logits of pass 0's strong rows."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib as L
from ..core.hooks import FreeMatchThresholdingHook, PseudoLabelingHook
from ..core.registry import ALGORITHMS
from .srflexmatch import SRFlexMatch
from .utils import SSL_Argument, str2bool


@ALGORITHMS.register("srfreematch")
class SRFreeMatch(SRFlexMatch):
    def _init_algorithm(self, args):
        self.init(T=args.T, hard_label=args.hard_label, ema_p=args.ema_p, use_quantile=args.use_quantile, clip_thresh=args.clip_thresh)
        self.lambda_e = args.ent_loss_ratio

    def init(self, T, hard_label=True, ema_p=0.999, use_quantile=True, clip_thresh=False):
        self.T, self.use_hard_label, self.ema_p, self.use_quantile, self.clip_thresh = T, hard_label, ema_p, use_quantile, clip_thresh

    def set_hooks(self):
        self.register_hook(PseudoLabelingHook(), "PseudoLabelingHook")
        self.register_hook(FreeMatchThresholdingHook(num_classes=self.num_classes, momentum=self.args.ema_p,
                                                     device=f"cuda:{self.gpu}" if torch.cuda.is_available() else "cpu"), "MaskingHook")
        super(SRFlexMatch, self).set_hooks()

    def _mask_and_pseudo(self, logits_w, idx_ulb, first_pass=True):
        mask = self.call_hook("masking", "MaskingHook", logits_x_ulb=logits_w, softmax_x_ulb=True, pseudo_from_probs=not first_pass)
        return mask, self._last_pseudo[1]

    def _extra_loss(self, losses, mask, logits_s, dl_s):
        """total += lambda_e * entropy_loss(mask, logits_s, p_model, label_hist)  (srfreematch.py:213-219)."""
        h = self.hooks_dict["MaskingHook"]
        ls = logits_s if logits_s.stride(-1) == 1 else logits_s.contiguous()
        a = L.FreeMatchEntropyArgs(B=ls.shape[0], num_classes=ls.shape[1], mask=mask.data_ptr(), logits_s=ls.data_ptr(), ld_logits=ls.stride(0),
                                   p_model=h.p_model.data_ptr(), label_hist=h.label_hist.data_ptr(), lambda_e=float(self.lambda_e),
                                   losses=losses.data_ptr(), dlogits_s=dl_s.data_ptr(), ld_dlogits=dl_s.stride(0), accumulate=1)
        L.check(L.load().srw_freematch_entropy(C.byref(a), L.stream_ptr()), "srw_freematch_entropy")

    def train_step(self, x_lb, y_lb, x_ulb_w, x_ulb_s):
        if not (torch.is_grad_enabled() and hasattr(self._net(), "forward_native")):
            raise RuntimeError("SRFreeMatch.train_step runs the native eager-backward step only (grad mode on, semireward_b200 ViT)")
        return self._train_step_eager(x_lb, y_lb, None, x_ulb_w, x_ulb_s)

    def get_save_dict(self):
        d = super(SRFlexMatch, self).get_save_dict()
        h = self.hooks_dict["MaskingHook"]
        d["p_model"], d["time_p"], d["label_hist"] = h.p_model.cpu(), h.time_p.reshape(()).cpu(), h.label_hist.cpu()   # 0-dim like the reference
        d["semireward"] = self._sr_save_dict()
        return d

    def load_model(self, load_path):
        ck = super(SRFlexMatch, self).load_model(load_path)
        h = self.hooks_dict["MaskingHook"]
        h.p_model, h.time_p, h.label_hist = ck["p_model"].cuda(self.gpu), ck["time_p"].cuda(self.gpu), ck["label_hist"].cuda(self.gpu)
        self._sr_load(ck)
        return ck

    @staticmethod
    def get_argument():
        return [SSL_Argument("--hard_label", str2bool, True), SSL_Argument("--T", float, 0.5), SSL_Argument("--p_cutoff", float, 0.95),
                SSL_Argument("--thresh_warmup", str2bool, True), SSL_Argument("--use_quantile", str2bool, False),
                SSL_Argument("--clip_thresh", str2bool, False), SSL_Argument("--ema_p", float, 0.999),
                SSL_Argument("--ent_loss_ratio", float, 0.01), SSL_Argument("--start_timing", int, 20000),
                SSL_Argument("--feature_dim", int, 384), SSL_Argument("--sr_lr", float, 0.0005), SSL_Argument("--N_k", int, 10),
                SSL_Argument("--sr_ema", str2bool, True), SSL_Argument("--sr_ema_m", float, 0.999)]
