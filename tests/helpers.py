"""Shared builders for the parity tests: the native algorithm and the oracle initialised from the same deterministic
fills (semireward_b200.detgen), stepping on the same deterministic batches."""
from __future__ import annotations

import functools

import torch


def small_cfg(algorithm="srflexmatch", **over):
    c = dict(algorithm=algorithm, net="vit_small_patch2_32", optim="AdamW", lr=5e-4, layer_decay=0.5, weight_decay=5e-4,
             num_train_iter=64, num_warmup_iter=0, start_timing=3, N_k=2, batch_size=8, uratio=1, num_classes=100,
             ulb_dest_len=64, feature_dim=384, sr_lr=5e-4, sr_ema=False, use_cat=True, amp=False, ema_m=0.0, gpu=0,
             thresh_warmup=True, p_cutoff=0.95)
    c.update(over)
    return c


def build_oracle(cfg: dict, depth: int, seed: int = 0, head_gain: float = 4.0):
    from oracle import ssl_oracle as O
    vc = O.ViTConfig(depth=depth, num_classes=cfg["num_classes"])
    sc = O.StepConfig(algorithm=cfg["algorithm"], num_classes=cfg["num_classes"], ulb_dest_len=cfg["ulb_dest_len"],
                      p_cutoff=cfg["p_cutoff"], thresh_warmup=cfg["thresh_warmup"], start_timing=cfg["start_timing"], N_k=cfg["N_k"],
                      num_train_iter=cfg["num_train_iter"], num_warmup_iter=cfg["num_warmup_iter"], lr=cfg["lr"],
                      weight_decay=cfg["weight_decay"], layer_decay=cfg["layer_decay"], sr_lr=cfg["sr_lr"], feature_dim=cfg["feature_dim"],
                      ema_p=cfg.get("ema_p", 0.999), use_quantile=cfg.get("use_quantile", True), clip_thresh=cfg.get("clip_thresh", False),
                      n_sigma=cfg.get("n_sigma", 2), lambda_e=cfg.get("ent_loss_ratio", 0.001), unsup_warm_up=cfg.get("unsup_warm_up", 0.4))
    return O.build_det_oracle(vc, sc, seed=seed, head_gain=head_gain)


def build_native(cfg: dict, depth: int, seed: int = 0, head_gain: float = 4.0):
    import semireward_b200 as S
    from semireward_b200 import detgen
    args = S.get_config(cfg)
    builder = functools.partial(S.get_net_builder(args.net, False), depth=depth, drop_path_rate=0.0)
    alg = S.get_algorithm(args, builder, None, None)
    with torch.no_grad():
        for prefix, mod in (("", alg.model), ("rewarder.", alg.rewarder), ("generator.", alg.generator)):
            for n, p in mod.named_parameters():
                p.copy_(torch.from_numpy(detgen.fill_param(prefix + n, p.shape, seed)))
                if prefix == "" and n == "head.weight":
                    p.mul_(head_gain)
    alg.model = alg.model.cuda(args.gpu).train()
    alg.rewarder = alg.rewarder.cuda(args.gpu)
    alg.generator = alg.generator.cuda(args.gpu)
    # the optimizer was built on the CPU parameters; .cuda() keeps the Parameter objects, so the groups stay valid
    return alg


def batch_tensors(cfg: dict, step: int, seed: int = 1):
    from semireward_b200 import detgen
    b = detgen.ssl_batch(cfg["batch_size"], cfg["uratio"], cfg["num_classes"], cfg["ulb_dest_len"], seed=seed, step=step)
    return {k: torch.from_numpy(v) for k, v in b.items()}
