"""SRFixMatch — FixMatch + SemiReward train step on the B200-native kernels, registered under the reference's name.

Follows semilearn/algorithms/srfixmatch/fixmatch.py: ctor :37-50, init :52-55, set_hooks :57-60, data_generator :62-94,
train_step :96-210, get_argument :212-226.  It is SRFlexMatch with the stateless FixedThresholdingHook
(mask = max_p >= p_cutoff, semilearn/algorithms/hooks/masking.py:42-57) and without idx_ulb / hook state in the
checkpoint; the step skeleton (backbone passes, Rewarder, SR online update, eager backward) is shared."""
from __future__ import annotations

import torch

from ..core.hooks import FixedThresholdingHook, PseudoLabelingHook
from ..core.registry import ALGORITHMS
from .srflexmatch import SRFlexMatch
from .utils import SSL_Argument, str2bool


@ALGORITHMS.register("srfixmatch")
class SRFixMatch(SRFlexMatch):
    def _init_algorithm(self, args):
        self.init(T=args.T, p_cutoff=args.p_cutoff, hard_label=args.hard_label)

    def init(self, T, p_cutoff, hard_label=True):
        self.T, self.p_cutoff, self.use_hard_label = T, p_cutoff, hard_label

    def set_hooks(self):
        self.register_hook(PseudoLabelingHook(), "PseudoLabelingHook")
        self.register_hook(FixedThresholdingHook(), "MaskingHook")
        super(SRFlexMatch, self).set_hooks()

    def _mask_and_pseudo(self, logits_w, idx_ulb, first_pass=True):
        mask = self.call_hook("masking", "MaskingHook", logits_x_ulb=logits_w, softmax_x_ulb=True)
        return mask, self._last_pseudo[1]   # hard labels = argmax of the probabilities (fixmatch.py:135-139, :83-87)

    def train_step(self, x_lb, y_lb, x_ulb_w, x_ulb_s):
        if not (torch.is_grad_enabled() and hasattr(self._net(), "forward_native")):
            raise RuntimeError("SRFixMatch.train_step runs the native eager-backward step only (grad mode on, semireward_b200 ViT)")
        return self._train_step_eager(x_lb, y_lb, None, x_ulb_w, x_ulb_s)

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
        return [SSL_Argument("--hard_label", str2bool, True), SSL_Argument("--T", float, 0.5), SSL_Argument("--p_cutoff", float, 0.95),
                SSL_Argument("--start_timing", int, 20000), SSL_Argument("--feature_dim", int, 384), SSL_Argument("--sr_lr", float, 0.0005),
                SSL_Argument("--N_k", int, 10), SSL_Argument("--sr_ema", str2bool, True), SSL_Argument("--sr_ema_m", float, 0.999)]
