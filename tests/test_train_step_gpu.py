"""-m gpu: SRFlexMatch.train_step + ParamUpdateHook on the native path vs the oracle (CPU restatement of the reference,
pinned against the live reference by tests/golden) on identical weights and batches, across stage 1, the gap step,
stage 2 with and without an SR update.

Gates (BASELINE.json north_star): pseudo-labels / mask / mask2 / selected_label bit-exact; logits, losses within 1e-3."""
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
        tap.update({n: p.grad.detach().clone() for n, p in alg.model.named_parameters()})
        return orig(*a, **k)

    alg.optimizer.step = step
    return tap


def _resync(alg, orc):
    """Copy the oracle's parameters into the native modules so every step is compared from identical state (Adam turns
    noise-level gradient entries, e.g. the mathematically-zero key-bias gradient, into +-lr moves, so free-running
    trajectories separate by O(lr) in those entries after one step on ANY two fp32 implementations)."""
    with torch.no_grad():
        for n, p in alg.model.named_parameters():
            p.copy_(orc.p[n].detach())
        for n, p in alg.rewarder.named_parameters():
            p.copy_(orc.rp[n].detach())


@pytest.mark.parametrize("depth,steps,over,resync", [
    (2, 8, {}, True),
    (2, 6, dict(thresh_warmup=False, p_cutoff=0.6), True),
    (12, 5, {}, True),
    (2, 8, {}, False),
])
def test_srflexmatch_steps_vs_oracle(depth, steps, over, resync):
    cfg = small_cfg(**over)
    orc = build_oracle(cfg, depth)
    alg = build_native(cfg, depth)
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
        tol = 1e-3 if resync else 5e-3
        for k_native, k_or in (("train/sup_loss", "sup_loss"), ("train/unsup_loss", "unsup_loss"), ("train/total_loss", "total_loss")):
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
