"""TEST INFRASTRUCTURE (build container only): drive the LIVE reference on CPU.

The reference (Westlake-AI/SemiReward, mounted read-only at /root/reference) is pure Python, so it can
be imported here to (a) validate the restatement in oracle/ssl_oracle.py and (b) generate the golden
vectors under tests/golden/.  It does NOT exist on the GPU box; nothing in tests -m gpu, smoke() or
bench.py imports this module.

Recipe = SURVEY.md Appendix A: import shims for packages the image lacks, CPU patches for the
hard-coded .cuda() calls, `dataset='synthetic'` so no loaders are built.  Nothing is copied from the
reference; it is only imported and called through its public API
(semilearn.get_algorithm / get_net_builder / AlgorithmBase.train_step / ParamUpdateHook).
"""
from __future__ import annotations

import inspect
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SRW_REFERENCE_ROOT", "/root/reference")
_loaded = False


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "semilearn"))


def load_reference():
    """Install shims + patches and import `semilearn` from the read-only reference tree."""
    global _loaded
    import torch
    import torch.nn as nn
    if _loaded:
        import semilearn
        return semilearn
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    import yaml
    # resolve transformers' lazies before a stub `timm` appears
    from transformers import BertModel, HubertModel, Wav2Vec2Model, Dinov2Model  # noqa: F401

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    stub("skimage").util = stub("skimage.util", montage=lambda *a, **k: None)
    stub("ruamel").yaml = stub("ruamel.yaml", load=yaml.load, Loader=yaml.Loader, dump=yaml.dump)
    stub("aim", Run=object)
    stub("matplotlib").pyplot = stub("matplotlib.pyplot")
    stub("progress").bar = stub("progress.bar", Bar=object)

    class DropPath(nn.Module):
        """timm semantics (not installed): per-sample Bernoulli(keep)/keep, identity in eval."""

        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            m = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep).div_(keep)
            return x * m

    stub("timm").models = stub("timm.models")
    sys.modules["timm.models"].layers = stub(
        "timm.models.layers", DropPath=DropPath,
        to_2tuple=lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v))

    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import semilearn  # noqa: E402
    for modname in ("srflexmatch.srflexmatch", "srfreematch.srfreematch", "srsoftmatch.srsoftmatch", "srfixmatch.fixmatch", "srpseudolabel.srpseudolabel"):
        mod = sys.modules.get(f"semilearn.algorithms.{modname}")
        if mod is not None:
            mod.send_model_cuda = lambda args, m, clip_batch=True: m
    _loaded = True
    return semilearn


DEFAULT_CFG = dict(
    algorithm="srflexmatch", net="vit_small_patch2_32", dataset="synthetic", num_classes=100, num_labels=200,
    batch_size=8, uratio=1, num_train_iter=204800, epoch=200, optim="AdamW", lr=5e-4, layer_decay=0.5,
    weight_decay=5e-4, use_cat=True, amp=False, ema_m=0.0, start_timing=20000, feature_dim=384, sr_lr=5e-4,
    N_k=10, sr_ema=False, sr_ema_m=0.99, gpu=0, distributed=False, use_wandb=False, use_aim=False,
    use_pretrain=False, num_warmup_iter=5120, ema_p=0.999, ent_loss_ratio=0.001, use_quantile=True,
    clip_thresh=False, img_size=32, ulb_dest_len=50000, drop_path=False)


def build_reference_algorithm(cfg: dict, net_kwargs: dict | None = None):
    """Build the reference algorithm object on CPU through its own public API.

    cfg: overrides of DEFAULT_CFG (keys are the reference's YAML keys; `ulb_dest_len` is set by hand because
    no dataset is built, `drop_path=False` zeroes DropPath for deterministic parity).
    net_kwargs: extra kwargs forwarded to the reference net builder (e.g. depth=2 for a small fixture).
    """
    import torch
    semilearn = load_reference()
    from semilearn.lighting.config import get_config
    c = dict(DEFAULT_CFG)
    c.update(cfg)
    ulb_dest_len = c.pop("ulb_dest_len")
    drop_path = c.pop("drop_path")
    args = get_config(c)
    args.ulb_dest_len = ulb_dest_len
    builder = semilearn.get_net_builder(args.net, False)
    nk = dict(net_kwargs or {})
    if not drop_path and args.net.startswith("vit"):
        nk["drop_path_rate"] = 0.0

    def net_builder(num_classes, pretrained=False, pretrained_path=None, **kw):
        kw = dict(kw)
        if args.net.startswith("wrn") and "depth" in nk:
            # the named builders fix depth 28 (wrn.py:160-171): a shallower fixture goes through the class directly
            from semilearn.nets.wrn.wrn import WideResNet
            return WideResNet(first_stride=1, num_classes=num_classes, depth=nk["depth"], widen_factor=dict(wrn_28_2=2, wrn_28_8=8)[args.net])
        if args.net.startswith("hubert"):
            # hubert.py:13 calls HubertModel.from_pretrained(name): hand it a randomly initialised HubertModel(HubertConfig(**hubert))
            import semilearn.nets.hubert.hubert as ref_hubert
            from transformers import HubertConfig, HubertModel
            hf = dict(nk.get("hubert", {}))
            orig = ref_hubert.HubertModel.from_pretrained
            ref_hubert.HubertModel.from_pretrained = classmethod(lambda cls, name, **k: HubertModel(HubertConfig(**hf)))
            try:
                m = builder(num_classes=num_classes, pretrained=False, pretrained_path=None)
            finally:
                ref_hubert.HubertModel.from_pretrained = orig
            if "dropout" in nk:
                m.dropout.p = nk["dropout"]
            return m
        if args.net.startswith("bert"):
            # bert.py:13 calls BertModel.from_pretrained(name): no hub access here, so hand it a randomly initialised BertModel of
            # the requested (small) configuration instead; `bert` = BertConfig overrides, `dropout` = the wrapper's own p
            import semilearn.nets.bert.bert as ref_bert
            from transformers import BertConfig, BertModel
            hf = dict(nk.get("bert", {}))
            orig = ref_bert.BertModel.from_pretrained
            ref_bert.BertModel.from_pretrained = classmethod(lambda cls, name, **k: BertModel(BertConfig(**hf)))
            try:
                m = builder(num_classes=num_classes, pretrained=False, pretrained_path=None)
            finally:
                ref_bert.BertModel.from_pretrained = orig
            if "dropout" in nk:
                m.dropout.p = nk["dropout"]
            return m
        for k, v in nk.items():
            kw[k] = v
        if "drop_path_rate" in kw:
            # the named builders pass drop_path_rate themselves -> go through the class directly
            from semilearn.nets.vit.vit import VisionTransformer
            base = dict(vit_small_patch2_32=dict(img_size=32, patch_size=2, embed_dim=384, depth=12, num_heads=6),
                        vit_tiny_patch2_32=dict(img_size=32, patch_size=2, embed_dim=192, depth=12, num_heads=3),
                        vit_base_patch16_224=dict(patch_size=16, embed_dim=768, depth=12, num_heads=12),
                        vit_base_patch16_96=dict(img_size=96, patch_size=16, embed_dim=768, depth=12, num_heads=12),
                        )[args.net]
            base.update(kw)
            return VisionTransformer(num_classes=num_classes, **base)
        return builder(num_classes=num_classes, pretrained=False, pretrained_path=None, **kw)

    torch.manual_seed(0)
    alg = semilearn.get_algorithm(args, net_builder, None, None)
    alg.model.train()
    return alg


def load_det_weights(alg, seed: int = 0, head_gain: float = 1.0):
    """Overwrite backbone / Rewarder / Generator parameters with semireward_b200.detgen fills."""
    import torch
    from semireward_b200 import detgen
    with torch.no_grad():
        for prefix, mod in (("", alg.model), ("rewarder.", alg.rewarder), ("generator.", alg.generator)):
            for n, p in mod.named_parameters():
                p.copy_(torch.from_numpy(detgen.fill_param(prefix + n, p.shape, seed)))
                if prefix == "" and n in ("head.weight", "classifier.2.weight", "classifier.weight"):
                    p.mul_(head_gain)
        alg.ema_model.load_state_dict(alg.model.state_dict())
    # optimizers hold references to the same Parameter objects -> nothing to rebuild


def run_reference_step(alg, batch: dict, it: int, do_update: bool = True):
    """alg.it = it; train_step; ParamUpdateHook.after_train_step (backward+optimizer+scheduler+zero_grad)."""
    import torch
    alg.it = it
    b = {k: (torch.from_numpy(v) if not isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    b = {k: v for k, v in b.items() if k in inspect.signature(alg.train_step).parameters}
    alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**b))
    if do_update:
        alg.hooks_dict["ParamUpdateHook"].after_train_step(alg)
    return alg.out_dict, alg.log_dict
