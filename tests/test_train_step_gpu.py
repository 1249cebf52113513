"""-m gpu: SRFlexMatch.train_step + ParamUpdateHook on the native path vs the oracle (CPU restatement of the reference,
pinned against the live reference by tests/golden) on identical weights and batches, across stage 1, the gap step,
stage 2 with and without an SR update.

Gates (BASELINE.json north_star): pseudo-labels / mask / mask2 / selected_label bit-exact; logits, losses within 1e-3."""
import numpy as np
import pytest
import torch

from helpers import batch_tensors, build_native, build_oracle, small_cfg

pytestmark = pytest.mark.gpu


def _grad_tap(alg):
    """Record the gradients the optimizer is about to consume (ParamUpdateHook zeroes them afterwards)."""
    tap = {}
    orig = alg.optimizer.step

    def step(*a, **k):
        tap.clear()
        tap.update({(n[7:] if n.startswith('module.') else n): p.grad.detach().clone() for n, p in alg.model.named_parameters() if p.grad is not None})   # BERT's pooler has none; a data-parallel wrapper prefixes 'module.'
        return orig(*a, **k)

    alg.optimizer.step = step
    return tap


def _mask2_tie(rec):
    """mask2 = reward >= reward.mean() (srflexmatch.py:100-101) is a rounding coin-flip when samples sit exactly at the mean:
    samples that share a pseudo-label get bit-identical rewards, and whether r+r+...+r over B equals B*r depends on the
    reduction order (torch CPU sums sequentially below its vector width, torch CUDA and srw_ssl_loss sum pairwise, which is
    exact for power-of-two batches).  Steps where that happens are not compared on mask2-dependent quantities."""
    if "dg_reward" not in rec:
        return False
    r = rec["dg_reward"].flatten().double()
    return bool(((r - r.mean()).abs() <= 4e-7 * r.mean().abs()).any())


def _resync(alg, orc):
    """Copy the oracle's parameters into the native modules so every step is compared from identical state (Adam turns
    noise-level gradient entries, e.g. the mathematically-zero key-bias gradient, into +-lr moves, so free-running
    trajectories separate by O(lr) in those entries after one step on ANY two fp32 implementations)."""
    with torch.no_grad():
        for n, p in alg.model.named_parameters():
            p.copy_(orc.p[n].detach())
        for n, p in alg.rewarder.named_parameters():
            p.copy_(orc.rp[n].detach())


@pytest.mark.parametrize("depth,steps,over,resync,eager", [
    (2, 8, {}, True, True),
    (2, 6, dict(thresh_warmup=False, p_cutoff=0.6), True, True),
    (12, 5, {}, True, True),
    (2, 8, {}, False, True),
    (2, 8, {}, True, False),    # the autograd route (_VitFunction / _SSLLoss) instead of the eager backward
    (2, 5, dict(batch_size=40, uratio=2, ulb_dest_len=512), True, True),   # 200 images per step (uratio 2): multi-wave GEMMs, ragged tiles
])
def test_srflexmatch_steps_vs_oracle(depth, steps, over, resync, eager):
    cfg = small_cfg(**over)
    orc = build_oracle(cfg, depth)
    alg = build_native(cfg, depth)
    alg.eager_backward = eager
    hook = alg.hooks_dict["MaskingHook"]
    tap = _grad_tap(alg)
    seen_mask_values = set()
    for it in range(steps):
        batch = batch_tensors(cfg, it)
        rec = orc.train_step(dict(batch), it)
        ref_grads = orc.param_update()
        alg.it = it
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**batch))
        alg.call_hook("after_train_step")
        torch.cuda.synchronize()
        ld = alg.log_dict
        for k_native, k_or in (("train/sup_loss", "sup_loss"), ("train/unsup_loss", "unsup_loss"), ("train/total_loss", "total_loss")):
            # gate: 1e-3 abs from identical state; the free-running trajectory (no resync) separates chaotically through Adam
            # (see _resync), so it is only held to 5e-3 relative (losses are ~1..16 here, head_gain 4)
            tol = 1e-3 if resync else 5e-3 * max(1.0, abs(float(rec[k_or])))
            assert abs(ld[k_native] - float(rec[k_or])) < tol, f"it {it} {k_or}: {ld[k_native]} vs {float(rec[k_or])}"
        assert abs(ld["train/util_ratio"] - float(rec["util_ratio"])) < 1e-6
        # bit-exact integer / mask state
        assert torch.equal(alg._last_pseudo_label.cpu(), rec["pseudo"]), f"it {it}: pseudo labels differ"
        assert torch.equal(alg._last_mask.cpu(), rec["mask"]), f"it {it}: mask differs"
        if "dg_mask2" in rec:
            assert torch.equal(alg._last_mask2.cpu(), rec["dg_mask2"]), f"it {it}: mask2 differs"
            seen_mask_values.update(rec["dg_mask2"].tolist())
        assert torch.equal(hook.selected_label.cpu(), orc.hook.selected_label), f"it {it}: selected_label differs"
        assert torch.equal(hook.classwise_acc.cpu(), orc.hook.classwise_acc), f"it {it}: classwise_acc differs"
        seen_mask_values.update(rec["mask"].tolist())
        # gradients (relative to each tensor's largest entry) and the parameters after AdamW where Adam is well conditioned
        worst_g = worst_p = worst_r = 0.0
        for n, p in alg.model.named_parameters():
            gr = ref_grads[n]
            sc = gr.abs().max().item()
            eg = (tap[n].cpu() - gr).abs().max().item() / max(sc, 1e-20)
            worst_g = max(worst_g, eg)
            well = gr.abs() > max(1e-2 * sc, 1e-6)
            dp = (p.detach().cpu() - orc.p[n].detach()).abs()
            if well.any():
                worst_p = max(worst_p, dp[well].max().item())
        for n, p in alg.rewarder.named_parameters():
            dp = (p.detach().cpu() - orc.rp[n].detach()).abs()
            gr = rec.get("sr_grads", {}).get(n)
            if gr is not None:   # an SR update happened this step: Adam is only well conditioned away from noise-level entries
                well = gr.abs() > max(1e-2 * gr.abs().max().item(), 1e-6)   # (e.g. d/d cross_attention_fc.bias is mathematically zero)
                dp = dp[well] if well.any() else dp[:0]
            if dp.numel():
                worst_r = max(worst_r, dp.max().item())
        print(f"it {it}: total {ld['train/total_loss']:.5f} (oracle {float(rec['total_loss']):.5f}) util {ld['train/util_ratio']:.3f} "
              f"grad rel err {worst_g:.2e} param diff (well-conditioned) {worst_p:.2e} rewarder maxdiff {worst_r:.2e}")
        if resync:
            assert worst_g < 1e-3, f"it {it}: gradient error {worst_g}"
            assert worst_p < 2e-5, f"it {it}: AdamW result differs {worst_p}"
            assert worst_r < 2e-5, f"it {it}: rewarder parameters drifted {worst_r}"
            _resync(alg, orc)
    print("mask values seen:", sorted(seen_mask_values))


@pytest.mark.parametrize("algorithm", ["srflexmatch", "srfreematch"])
def test_stochastic_stage2_batched_matches_sequential_passes(algorithm):
    """DropPath on: stage 2 needs 1 + K backbone passes per step with fresh DropPath draws (like the reference), and two
    graphs carry gradient (pass 0 -> sup loss [+ FreeMatch entropy], last pass -> unsup loss).  The batched route (all used
    rows in ONE forward + ONE backward) must reproduce the sequential passes from the same RNG state: identical masks and
    pseudo-labels, losses to fp32 rounding, gradients to 1e-4 relative (the wgrad split-K partition differs)."""
    import functools
    import semireward_b200 as S
    from semireward_b200 import detgen
    cfg = small_cfg(algorithm=algorithm, num_train_iter=8, start_timing=1, ent_loss_ratio=0.05, use_quantile=True)   # it=3: K = 8
    results = []
    for batched in (True, False):
        args = S.get_config(cfg)
        builder = functools.partial(S.get_net_builder(args.net, False), depth=2, drop_path_rate=0.3)
        alg = S.get_algorithm(args, builder, None, None)
        with torch.no_grad():
            for prefix, mod in (("", alg.model), ("rewarder.", alg.rewarder), ("generator.", alg.generator)):
                for n, p in mod.named_parameters():
                    p.copy_(torch.from_numpy(detgen.fill_param(prefix + n, p.shape, 0)))
        alg.model = alg.model.cuda(args.gpu).train()
        alg.rewarder, alg.generator = alg.rewarder.cuda(args.gpu), alg.generator.cuda(args.gpu)
        alg.batch_stochastic_passes = batched
        tap = _grad_tap(alg)
        torch.manual_seed(123)
        rec = []
        for it in (2, 3, 4):
            alg.it = it
            alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**batch_tensors(cfg, it)))
            alg.call_hook("after_train_step")
            torch.cuda.synchronize()
            rec.append((dict(alg.log_dict), alg._last_mask.clone(), alg._last_mask2.clone(), alg._last_pseudo_label.clone(),
                        {n: g.clone() for n, g in tap.items()}))
        results.append(rec)
    for (ld_a, m_a, m2_a, p_a, g_a), (ld_b, m_b, m2_b, p_b, g_b) in zip(*results):
        for k in ("train/sup_loss", "train/unsup_loss", "train/total_loss", "train/util_ratio"):
            assert abs(ld_a[k] - ld_b[k]) < 2e-5 * max(1.0, abs(ld_b[k])), (k, ld_a[k], ld_b[k])
        assert torch.equal(m_a, m_b) and torch.equal(m2_a, m2_b) and torch.equal(p_a, p_b)
        for n in g_a:
            sc = g_b[n].abs().max().item()
            assert (g_a[n] - g_b[n]).abs().max().item() <= 1e-4 * max(sc, 1e-20), n


@pytest.mark.parametrize("algorithm,over", [
    ("srfreematch", dict(use_quantile=True, clip_thresh=False, ent_loss_ratio=0.05)),
    ("srfreematch", dict(use_quantile=False, clip_thresh=True, ent_loss_ratio=0.05)),
    ("srsoftmatch", dict(n_sigma=2)),
    ("srfixmatch", dict(p_cutoff=0.5)),
    ("srpseudolabel", dict(p_cutoff=0.5, unsup_warm_up=0.05)),
])
def test_srfreematch_srsoftmatch_steps_vs_oracle(algorithm, over):
    """SRFreeMatch / SRSoftMatch native steps (stage 1, the gap step, stage 2 with and without an SR update) against the
    oracle from identical parameters: hard pseudo-labels and 0/1 masks bit-exact, SoftMatch's soft weights and the EMA state
    within 1e-4 (they integrate probabilities of logits that agree to ~4e-4), losses within 1e-3, gradients within 1e-3 relative."""
    cfg = small_cfg(algorithm=algorithm, ema_p=0.9, **over)   # momentum 0.9: the state moves visibly within 8 steps
    hg = 2.0 if algorithm in ("srfixmatch", "srpseudolabel") else 4.0   # mixed masks at p_cutoff 0.5 (as in the golden cases)
    orc = build_oracle(cfg, 2, head_gain=hg)
    alg = build_native(cfg, 2, head_gain=hg)
    tap = _grad_tap(alg)
    hook = alg.hooks_dict["MaskingHook"]
    for it in range(8):
        batch = batch_tensors(cfg, it)
        rec = orc.train_step(dict(batch), it)
        ref_grads = orc.param_update()
        alg.it = it
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**batch))
        alg.call_hook("after_train_step")
        torch.cuda.synchronize()
        ld = alg.log_dict
        tie = _mask2_tie(rec)
        for k_native, k_or in (("train/sup_loss", "sup_loss"), ("train/unsup_loss", "unsup_loss"), ("train/total_loss", "total_loss")):
            if tie and k_or != "sup_loss":
                continue
            assert abs(ld[k_native] - float(rec[k_or])) < 1e-3, f"it {it} {k_or}: {ld[k_native]} vs {float(rec[k_or])}"
        assert abs(ld["train/util_ratio"] - float(rec["util_ratio"])) < 1e-5
        assert torch.equal(alg._last_pseudo_label.cpu(), rec["pseudo"]), f"it {it}: pseudo labels differ"
        if algorithm in ("srfixmatch", "srpseudolabel"):
            assert torch.equal(alg._last_mask.cpu(), rec["mask"]), f"it {it}: mask differs"
        elif algorithm == "srfreematch":
            assert torch.equal(alg._last_mask.cpu(), rec["mask"]), f"it {it}: mask differs"
            # the EMA inputs are probabilities of logits that agree to ~4e-4 (bf16x3 vs fp32): state agrees to ~1e-5
            assert (hook.p_model.cpu() - orc.hook.p_model).abs().max().item() < 1e-4
            assert (hook.label_hist.cpu() - orc.hook.label_hist).abs().max().item() < 1e-6   # integer histogram EMA
            assert abs(hook.time_p.item() - float(orc.hook.time_p)) < 1e-4
        else:
            assert (alg._last_mask.cpu() - rec["mask"]).abs().max().item() < 1e-4, f"it {it}: weights differ"
            assert abs(hook.prob_max_mu_t.item() - float(orc.hook.prob_max_mu_t)) < 1e-4
            assert abs(hook.prob_max_var_t.item() - float(orc.hook.prob_max_var_t)) < 1e-4
            assert (alg.hooks_dict["DistAlignHook"].p_model.cpu() - orc.da.p_model).abs().max().item() < 1e-4
        if tie:
            print(f"{algorithm} it {it}: rewards tie with their mean -> mask2-dependent checks skipped (native mask2 {alg._last_mask2.tolist()})")
            _resync(alg, orc)
            continue
        if "dg_mask2" in rec:
            assert torch.equal(alg._last_mask2.cpu(), rec["dg_mask2"]), f"it {it}: mask2 differs"
        worst_g = 0.0
        for n, p in alg.model.named_parameters():
            gr = ref_grads[n]
            worst_g = max(worst_g, (tap[n].cpu() - gr).abs().max().item() / max(gr.abs().max().item(), 1e-20))
        print(f"{algorithm} it {it}: total {ld['train/total_loss']:.5f} (oracle {float(rec['total_loss']):.5f}) util {ld['train/util_ratio']:.3f} grad rel err {worst_g:.2e}")
        assert worst_g < 1e-3, f"it {it}: gradient error {worst_g}"
        _resync(alg, orc)


def test_config3_shape_srfreematch_vit_base_patch16_224():
    """BASELINE configs[2] at parity-test size: SRFreeMatch on vit_base_patch16_224 (N = 197 tokens, D = 768, 12 heads,
    patch-embed K = 768), 1000 classes, use_quantile — 2 blocks, batch 2+2+2 — against the oracle (stage 1, gap, stage 2)."""
    import functools
    import semireward_b200 as S
    from oracle import ssl_oracle as O
    from semireward_b200 import detgen
    cfg = small_cfg(algorithm="srfreematch", net="vit_base_patch16_224", num_classes=1000, batch_size=2, feature_dim=768, img_size=224,
                    use_quantile=True, clip_thresh=False, ent_loss_ratio=0.05, ema_p=0.9, start_timing=2, num_train_iter=16)
    depth = 2
    vc = O.ViTConfig(img_size=224, patch_size=16, embed_dim=768, depth=depth, num_heads=12, num_classes=1000)
    sc = O.StepConfig(algorithm="srfreematch", num_classes=1000, ulb_dest_len=cfg["ulb_dest_len"], start_timing=2, N_k=cfg["N_k"],
                      num_train_iter=16, num_warmup_iter=0, lr=cfg["lr"], weight_decay=cfg["weight_decay"], layer_decay=cfg["layer_decay"],
                      sr_lr=cfg["sr_lr"], feature_dim=768, ema_p=0.9, use_quantile=True, clip_thresh=False, lambda_e=0.05)
    orc = O.build_det_oracle(vc, sc, seed=0, head_gain=4.0)
    args = S.get_config(cfg)
    alg = S.get_algorithm(args, functools.partial(S.get_net_builder(args.net, False), depth=depth, drop_path_rate=0.0), None, None)
    with torch.no_grad():
        for prefix, mod in (("", alg.model), ("rewarder.", alg.rewarder), ("generator.", alg.generator)):
            for n, p in mod.named_parameters():
                p.copy_(torch.from_numpy(detgen.fill_param(prefix + n, p.shape, 0)))
                if prefix == "" and n == "head.weight":
                    p.mul_(4.0)
    alg.model = alg.model.cuda(args.gpu).train()
    alg.rewarder, alg.generator = alg.rewarder.cuda(args.gpu), alg.generator.cuda(args.gpu)
    tap = _grad_tap(alg)
    for it in range(4):
        b = detgen.ssl_batch(2, 1, 1000, cfg["ulb_dest_len"], img_size=224, seed=1, step=it)
        batch = {k: torch.from_numpy(v) for k, v in b.items()}
        rec = orc.train_step(dict(batch), it)
        ref_grads = orc.param_update()
        alg.it = it
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**batch))
        alg.call_hook("after_train_step")
        torch.cuda.synchronize()
        ld = alg.log_dict
        assert torch.equal(alg._last_pseudo_label.cpu(), rec["pseudo"]) and torch.equal(alg._last_mask.cpu(), rec["mask"])
        assert abs(ld["train/sup_loss"] - float(rec["sup_loss"])) < 1e-3
        if _mask2_tie(rec):
            _resync(alg, orc)
            continue
        for k_native, k_or in (("train/unsup_loss", "unsup_loss"), ("train/total_loss", "total_loss")):
            assert abs(ld[k_native] - float(rec[k_or])) < 1e-3, f"it {it} {k_or}: {ld[k_native]} vs {float(rec[k_or])}"
        worst_g = 0.0
        for n, p in alg.model.named_parameters():
            gr = ref_grads[n]
            worst_g = max(worst_g, (tap[n].cpu() - gr).abs().max().item() / max(gr.abs().max().item(), 1e-20))
        print(f"vit_base_patch16_224 it {it}: total {ld['train/total_loss']:.5f} (oracle {float(rec['total_loss']):.5f}) grad rel err {worst_g:.2e}")
        assert worst_g < 1e-3
        _resync(alg, orc)


def test_srpseudolabel_stochastic_stage2_runs():
    """DropPath on, stage 2: SRPseudoLabel runs [x_lb | x_ulb_w (pass K) | x_ulb_w (pass 0)] as one forward (see the module
    docstring).  No sequential native twin exists to compare with, so this checks shapes, finiteness and that pass 0 and
    pass K really see different DropPath draws (their logits differ) while gradients reach every parameter."""
    import functools
    import semireward_b200 as S
    from semireward_b200 import detgen
    cfg = small_cfg(algorithm="srpseudolabel", num_train_iter=8, start_timing=1, p_cutoff=0.5, unsup_warm_up=0.05)
    args = S.get_config(cfg)
    alg = S.get_algorithm(args, functools.partial(S.get_net_builder(args.net, False), depth=2, drop_path_rate=0.3), None, None)
    with torch.no_grad():
        for prefix, mod in (("", alg.model), ("rewarder.", alg.rewarder), ("generator.", alg.generator)):
            for n, p in mod.named_parameters():
                p.copy_(torch.from_numpy(detgen.fill_param(prefix + n, p.shape, 0)))
    alg.model = alg.model.cuda(args.gpu).train()
    alg.rewarder, alg.generator = alg.rewarder.cuda(args.gpu), alg.generator.cuda(args.gpu)
    tap = _grad_tap(alg)
    torch.manual_seed(5)
    for it in (2, 3, 4):
        alg.it = it
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**batch_tensors(cfg, it)))
        alg.call_hook("after_train_step")
        torch.cuda.synchronize()
        assert all(np.isfinite(v) for v in alg.log_dict.values()), alg.log_dict
        assert alg._last_mask.shape == (8,) and alg._last_mask2.shape == (8,)
        assert all(torch.isfinite(g).all() and g.abs().max() > 0 for n, g in tap.items() if not n.endswith("attn.qkv.bias")), it
