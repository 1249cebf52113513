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


def build_oracle(cfg: dict, depth: int, seed: int = 0, head_gain: float = 4.0, drop_path_rate: float = 0.0):
    from oracle import ssl_oracle as O
    vc = O.ViTConfig(depth=depth, num_classes=cfg["num_classes"], drop_path_rate=drop_path_rate)
    sc = O.StepConfig(algorithm=cfg["algorithm"], num_classes=cfg["num_classes"], ulb_dest_len=cfg["ulb_dest_len"],
                      p_cutoff=cfg["p_cutoff"], thresh_warmup=cfg["thresh_warmup"], start_timing=cfg["start_timing"], N_k=cfg["N_k"],
                      num_train_iter=cfg["num_train_iter"], num_warmup_iter=cfg["num_warmup_iter"], lr=cfg["lr"],
                      weight_decay=cfg["weight_decay"], layer_decay=cfg["layer_decay"], sr_lr=cfg["sr_lr"], feature_dim=cfg["feature_dim"],
                      ema_p=cfg.get("ema_p", 0.999), use_quantile=cfg.get("use_quantile", True), clip_thresh=cfg.get("clip_thresh", False),
                      n_sigma=cfg.get("n_sigma", 2), lambda_e=cfg.get("ent_loss_ratio", 0.001), unsup_warm_up=cfg.get("unsup_warm_up", 0.4))
    return O.build_det_oracle(vc, sc, seed=seed, head_gain=head_gain)


def build_native(cfg: dict, depth: int, seed: int = 0, head_gain: float = 4.0, drop_path_rate: float = 0.0):
    import semireward_b200 as S
    from semireward_b200 import detgen
    args = S.get_config(cfg)
    builder = functools.partial(S.get_net_builder(args.net, False), depth=depth, drop_path_rate=drop_path_rate)
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


class SharedDropPath:
    """One stream of DropPath multipliers for BOTH sides of a parity test.  The oracle draws `[depth, 2, rows]` per backbone call
    in its row order (lb, weak, strong) — this object draws them (CPU generator), remembers the sequence of one step, and
    replays it to the native net in the layout its two routes ask for: `[depth, 2, per]` per pass in engine row order
    (lb, strong, weak) for sequential passes, or `[depth, 2, (K + 1) * per]` = passes-major for the batched stage-2 route."""

    def __init__(self, vit_cfg, nl, nu, seed=1234):
        from oracle import ssl_oracle as O
        self.O, self.vc, self.nl, self.nu = O, vit_cfg, nl, nu
        self.gen = torch.Generator().manual_seed(seed)
        self.step_draws, self.native_cursor = [], 0
        self.perm = torch.cat([torch.arange(nl), torch.arange(nl + nu, nl + 2 * nu), torch.arange(nl, nl + nu)])

    def install(self, orc, alg):
        O, this = self.O, self
        orig = O.draw_drop_path_masks

        def oracle_draw(cfg, batch, generator=None):
            m = orig(cfg, batch, this.gen)
            this.step_draws.append(m)
            return m
        O.draw_drop_path_masks = oracle_draw
        self._restore = lambda: setattr(O, "draw_drop_path_masks", orig)
        net = alg._net()
        per = self.nl + 2 * self.nu

        def native_draw(batch, device, out=None):
            if batch == per:
                m = this.step_draws[this.native_cursor][:, :, this.perm]
                this.native_cursor += 1
            else:   # batched stage 2: every pass of the step at once, [depth, 2, pass, engine rows]
                assert batch == len(this.step_draws) * per, (batch, len(this.step_draws), per)
                m = torch.stack([d[:, :, this.perm] for d in this.step_draws], dim=2).reshape(this.step_draws[0].shape[0], 2, -1)
            m = m.to(device).contiguous()
            if out is not None:
                out.copy_(m)
                return out
            return m
        net._draw_drop_scale = native_draw

    def new_step(self):
        self.step_draws, self.native_cursor = [], 0

    def uninstall(self):
        self._restore()
