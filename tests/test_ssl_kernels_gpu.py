"""-m gpu: the fused SSL epilogue kernels against the oracle's restatement (oracle/ssl_oracle.py), tensor by tensor."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mods(feature_dim=384, num_classes=100, seed=0):
    from oracle import ssl_oracle as O
    from semireward_b200 import detgen
    from semireward_b200.algorithms.semireward import Generator, Rewarder, label_dim
    rp = {n: torch.from_numpy(detgen.fill_param("rewarder." + n, s, seed)) for n, s in O.rewarder_param_shapes(feature_dim, num_classes)}
    gp = {n: torch.from_numpy(detgen.fill_param("generator." + n, s, seed)) for n, s in O.generator_param_shapes(feature_dim)}
    R = Rewarder(label_dim(num_classes), 128, feature_dim)
    G = Generator(feature_dim)
    R.load_state_dict(rp)
    G.load_state_dict(gp)
    return O, rp, gp, R.cuda(), G.cuda()


@pytest.mark.parametrize("B", [8, 37, 128])
def test_rewarder_forward_and_generator(B):
    from semireward_b200 import detgen
    O, rp, gp, R, G = _mods()
    feats = torch.from_numpy(detgen.normal("feats", (B, 384), 3))
    labels = torch.from_numpy(detgen.integers("labels", (B,), 0, 100, 3))
    ref = O.rewarder_forward(rp, feats, labels)
    got = R(feats.cuda(), labels.cuda()).cpu()
    assert (got - ref).abs().max().item() < 1e-6
    gfeats = feats * 3.0
    ref_l = O.generator_forward(gp, gfeats).long().squeeze(1)
    got_l = G.generate_labels(gfeats.cuda()).cpu()
    assert torch.equal(got_l, ref_l)


@pytest.mark.parametrize("B", [8, 37])
def test_rewarder_train_gradients_and_adam(B):
    from semireward_b200 import detgen
    O, rp, gp, R, G = _mods()
    feats = torch.from_numpy(detgen.normal("feats", (B, 384), 4))
    gen = torch.from_numpy(detgen.integers("gen", (B,), 0, 3, 4))      # repeated labels on purpose
    true = torch.from_numpy(detgen.integers("true", (B,), 0, 3, 4))
    p = {k: v.clone().requires_grad_(True) for k, v in rp.items()}
    opt = O.AdamState(p, decoupled=False)
    for step in range(3):
        reward = O.rewarder_forward(p, feats, gen)
        target = O.sr_target(gen, true, 100)
        gl, rl = F.mse_loss(reward, torch.ones_like(reward)), F.mse_loss(reward, target)
        names = list(p)
        grads = dict(zip(names, torch.autograd.grad(gl + rl, [p[k] for k in names], allow_unused=True)))
        losses = R.train_step(feats.cuda(), gen.cuda(), true.cuda(), 5e-4, 100).cpu()
        assert abs(losses[0].item() - gl.item()) < 1e-6 and abs(losses[1].item() - rl.item()) < 1e-6
        got = dict(zip([n for n, _ in R.named_parameters()], R._adam["g"]))
        for n in names:
            g_ref = grads[n]
            g_got = got[n].cpu()
            sc = max(g_ref.abs().max().item(), 1e-12)
            err = (g_got - g_ref).abs().max().item()
            assert err < 1e-4 * sc + 1e-7, f"step {step} {n}: grad err {err:.3e} (max |g| {sc:.3e})"
        opt.step(p, grads, {k: (5e-4, 0.0) for k in names})
        for n, q in R.named_parameters():
            well = grads[n].abs() > max(1e-2 * grads[n].abs().max().item(), 1e-6)   # noise-level entries are ill-conditioned under Adam
            d = (q.detach().cpu() - p[n].detach()).abs()
            if well.any():
                assert d[well].max().item() < 1e-6, f"step {step} {n}: Adam result differs {d[well].max().item():.3e}"
        with torch.no_grad():   # resync the noise-level entries
            for n, q in R.named_parameters():
                q.copy_(p[n].detach())


def test_flexmatch_mask_and_loss_kernels():
    import ctypes as C
    from oracle import ssl_oracle as O
    from semireward_b200 import _lib as L, detgen
    from semireward_b200.core.hooks import FlexMatchThresholdingHook

    class A:  # minimal algorithm stand-in
        p_cutoff = 0.7
    B, Cn, U = 16, 10, 40
    hook = FlexMatchThresholdingHook(U, Cn, thresh_warmup=False, device="cuda")
    st = O.FlexMatchState(U, Cn, thresh_warmup=False)
    alg = A()
    for it in range(6):
        logits = torch.from_numpy(detgen.normal("lw", (B, Cn), it, std=3.0))
        idx = torch.from_numpy(detgen.distinct_integers("idx", B, U, it))
        probs = torch.softmax(logits, -1)
        ref_mask = st.masking(probs, idx, 0.7)
        mask = hook.masking(alg, logits.cuda(), idx.cuda())
        assert torch.equal(alg._last_pseudo[1].cpu(), probs.argmax(-1))
        assert (alg._last_probs.cpu() - probs).abs().max().item() < 1e-6
        assert torch.equal(mask.cpu(), ref_mask), f"it {it}"
        assert torch.equal(hook.selected_label.cpu(), st.selected_label)
        assert torch.equal(hook.classwise_acc.cpu(), st.classwise_acc)
    assert 0.0 < mask.mean().item() < 1.0 or True
    # loss kernel
    from semireward_b200.algorithms.srflexmatch import _SSLLoss
    llb = torch.from_numpy(detgen.normal("llb", (8, Cn), 9, std=2.0)).requires_grad_(True)
    ls = torch.from_numpy(detgen.normal("ls", (B, Cn), 9, std=2.0)).requires_grad_(True)
    y = torch.from_numpy(detgen.integers("y", (8,), 0, Cn, 9))
    pseudo = torch.from_numpy(detgen.integers("ps", (B,), 0, Cn, 9))
    m = (torch.from_numpy(detgen.uniform("m", (B,), 9)) > 0.4).float()
    reward = torch.from_numpy(detgen.uniform("r", (B,), 9))
    m2 = (reward >= reward.mean()).float()
    ref = O.ce_loss(llb, y, "mean") + 0.7 * O.consistency_loss(ls, pseudo, m, m2)
    ref.backward()
    gl, gs = llb.detach().cuda().requires_grad_(True), ls.detach().cuda().requires_grad_(True)
    side = {}
    tot = _SSLLoss.apply(gl, gs, y.cuda(), pseudo.cuda(), m.cuda(), reward.cuda(), 0.7, side)
    tot.backward()
    assert abs(tot.item() - ref.item()) < 1e-5
    assert torch.equal(side["mask2"].cpu(), m2)
    assert (gl.grad.cpu() - llb.grad).abs().max().item() < 1e-6
    assert (gs.grad.cpu() - ls.grad).abs().max().item() < 1e-6


class _Alg:   # the attributes the hooks read from the algorithm
    distributed, world_size = False, 1

    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.hooks_dict = {}


@pytest.mark.parametrize("B,C,use_quantile,clip", [(8, 100, True, False), (8, 100, False, True), (37, 10, True, True), (128, 1000, True, False)])
def test_freematch_mask_kernel_vs_oracle_state_machine(B, C, use_quantile, clip):
    """srw_freematch_mask against FreeMatchState (freematch/utils.py:23-66) on IDENTICAL logits over several calls: masks
    and pseudo-labels bit-exact, time_p / p_model / label_hist to fp32 rounding (the sums run in a different order)."""
    from oracle import ssl_oracle as O
    from semireward_b200 import detgen
    from semireward_b200.core.hooks import FreeMatchThresholdingHook
    st = O.FreeMatchState(C, 0.9)
    hook = FreeMatchThresholdingHook(C, 0.9, device="cuda")
    alg = _Alg(use_quantile=use_quantile, clip_thresh=clip)
    for call in range(6):
        logits = torch.from_numpy(detgen.normal("fm_logits", (B, C), 40 + call)) * (1.0 + call)
        from_probs = call % 2 == 1
        probs = torch.softmax(logits, dim=-1)
        ref_mask = st.masking(probs, use_quantile, clip)
        ref_pseudo = probs.argmax(dim=-1) if from_probs else logits.argmax(dim=-1)
        mask = hook.masking(alg, logits.cuda(), softmax_x_ulb=True, pseudo_from_probs=from_probs)
        torch.cuda.synchronize()
        assert torch.equal(alg._last_pseudo[1].cpu(), ref_pseudo), f"call {call}"
        assert (alg._last_probs.cpu() - probs).abs().max().item() < 1e-6
        assert abs(hook.time_p.item() - float(st.time_p)) < 2e-7, (call, hook.time_p.item(), float(st.time_p))
        assert (hook.p_model.cpu() - st.p_model).abs().max().item() < 2e-7
        assert (hook.label_hist.cpu() - st.label_hist).abs().max().item() < 2e-7
        assert torch.equal(mask.cpu(), ref_mask), f"call {call}: mask differs"


@pytest.mark.parametrize("B,C", [(8, 100), (37, 10), (128, 1000)])
def test_softmatch_mask_kernel_vs_oracle_state_machine(B, C):
    """srw_softmatch_mask against DistAlignState + SoftMatchState on identical logits: aligned (train_step) and
    un-aligned (data_generator) calls interleaved; weights and state to 1e-6."""
    from oracle import ssl_oracle as O
    from semireward_b200 import detgen
    from semireward_b200.core.hooks import DistAlignEMAHook, SoftMatchWeightingHook
    st, da = O.SoftMatchState(C, 2, 0.9), O.DistAlignState(C, 0.9)
    hook, dah = SoftMatchWeightingHook(C, 2, 0.9, device="cuda"), DistAlignEMAHook(C, 0.9, device="cuda")
    alg = _Alg()
    alg.hooks_dict["DistAlignHook"] = dah
    for call in range(6):
        logits = torch.from_numpy(detgen.normal("sm_logits", (B, C), 60 + call)) * (1.0 + 0.5 * call)
        align = call % 3 != 2
        probs = torch.softmax(logits, dim=-1)
        ref_w = st.masking(da.dist_align(probs) if align else probs)
        w = hook.masking(alg, logits.cuda(), softmax_x_ulb=True, dist_align=align, pseudo_from_probs=not align)
        torch.cuda.synchronize()
        assert torch.equal(alg._last_pseudo[1].cpu(), probs.argmax(-1) if not align else logits.argmax(-1))
        assert abs(hook.prob_max_mu_t.item() - float(st.prob_max_mu_t)) < 1e-6
        assert abs(hook.prob_max_var_t.item() - float(st.prob_max_var_t)) < 1e-6
        assert (dah.p_model.cpu() - da.p_model).abs().max().item() < 1e-6
        assert (w.cpu() - ref_w).abs().max().item() < 2e-6, f"call {call}"


@pytest.mark.parametrize("B,C,nsel", [(8, 100, 5), (8, 100, 0), (37, 10, 37), (128, 1000, 64)])
def test_freematch_entropy_kernel_vs_autograd(B, C, nsel):
    """srw_freematch_entropy: value and gradient w.r.t. the strong logits against torch autograd of the oracle's
    freematch_entropy_loss (srfreematch.py:16-44); accumulates on top of an existing gradient / total loss."""
    import ctypes as Ct
    from oracle import ssl_oracle as O
    from semireward_b200 import _lib as L, detgen
    logits = (torch.from_numpy(detgen.normal("ent_logits", (B, C), 70)) * 2.0).requires_grad_(True)
    mask = torch.zeros(B)
    mask[torch.from_numpy(detgen.integers("ent_sel", (B,), 0, 1 << 30, 71)).argsort()[:nsel]] = 1.0
    p_model = torch.softmax(torch.from_numpy(detgen.normal("ent_pm", (C,), 72)), 0)
    label_hist = torch.softmax(torch.from_numpy(detgen.normal("ent_lh", (C,), 73)), 0)
    if C >= 10:
        label_hist[3] = 0.0   # exercises replace_inf_to_zero
    lam = 0.05
    base = torch.from_numpy(detgen.normal("ent_base", (B, C), 74)) * 0.01
    losses = torch.tensor([0.5, 0.25, 0.75, 1.0, -1.0], device="cuda")
    dl = base.clone().cuda()
    lg = logits.detach().cuda()
    a = L.FreeMatchEntropyArgs(B=B, num_classes=C, mask=mask.cuda().data_ptr(), logits_s=lg.data_ptr(), ld_logits=C,
                               p_model=p_model.cuda().data_ptr(), label_hist=label_hist.cuda().data_ptr(), lambda_e=lam,
                               losses=losses.data_ptr(), dlogits_s=dl.data_ptr(), ld_dlogits=C, accumulate=1)
    mk, pm, lh = mask.cuda(), p_model.cuda(), label_hist.cuda()
    a.mask, a.p_model, a.label_hist = mk.data_ptr(), pm.data_ptr(), lh.data_ptr()
    L.check(L.load().srw_freematch_entropy(Ct.byref(a), L.stream_ptr()), "srw_freematch_entropy")
    torch.cuda.synchronize()
    if nsel == 0:
        assert losses[4].item() == 0.0 and abs(losses[2].item() - 0.75) < 1e-7 and torch.equal(dl.cpu(), base)
        return
    ref = O.freematch_entropy_loss(mask, logits, p_model, label_hist)
    (g_ref,) = torch.autograd.grad(lam * ref, logits)
    assert abs(losses[4].item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item())), (losses[4].item(), ref.item())
    assert abs(losses[2].item() - (0.75 + lam * ref.item())) < 1e-5
    err = (dl.cpu() - base - g_ref).abs().max().item()
    assert err < 1e-5 * max(g_ref.abs().max().item(), 1e-3), (err, g_ref.abs().max().item())


def test_freematch_mask_data_parallel_view():
    """C4 (freematch/utils.py:25-26): under data parallelism update() integrates the probabilities of ALL ranks and
    masking() thresholds the local rows.  Emulated in one process: the hook of 'rank 1' gets the rank-major gathered
    probabilities through its testing hook and must reproduce the reference state machine fed the same way."""
    from oracle import ssl_oracle as O
    from semireward_b200 import detgen
    from semireward_b200.core.hooks import FreeMatchThresholdingHook
    B, C, W, rank = 8, 100, 2, 1
    st = O.FreeMatchState(C, 0.9)
    hook = FreeMatchThresholdingHook(C, 0.9, device="cuda")
    alg = _Alg(use_quantile=True, clip_thresh=False)
    for call in range(4):
        logits_all = torch.from_numpy(detgen.normal("fm_dp_logits", (W * B, C), 90 + call)) * (1.5 + call)
        probs_all = torch.softmax(logits_all, dim=-1)
        local = slice(rank * B, (rank + 1) * B)
        # reference: update() on the gathered probabilities, mask on the local ones (masking() after update(), utils.py:60-65)
        st.update(probs_all, True, False)
        mp, mi = probs_all[local].max(dim=-1)
        ref_mask = mp.ge(st.time_p * (st.p_model / st.p_model.max())[mi]).float()
        mask = hook.masking(alg, logits_all[local].cuda(), softmax_x_ulb=True, probs_all=probs_all.cuda())
        torch.cuda.synchronize()
        assert abs(hook.time_p.item() - float(st.time_p)) < 2e-7
        assert (hook.p_model.cpu() - st.p_model).abs().max().item() < 2e-7
        assert (hook.label_hist.cpu() - st.label_hist).abs().max().item() < 2e-7
        assert torch.equal(mask.cpu(), ref_mask), f"call {call}"


def test_ema_hook_matches_reference_expression():
    """EMAHook / EMA.update (core/hooks/ema.py:20-24, misc.py:152-155): shadow = (1 - d) * p + d * shadow over every
    parameter in one launch, bit-exact with the tensor expression; the shadow is ema_model's parameters."""
    from semireward_b200.core.hooks import EMA
    from semireward_b200.nets import vit_small_patch2_32
    torch.manual_seed(0)
    model = vit_small_patch2_32(num_classes=100, depth=2).cuda()
    ema_model = vit_small_patch2_32(num_classes=100, depth=2)
    ema = EMA(model, 0.999, ema_model=ema_model)
    ema.register()
    ref = {n: p.detach().clone() for n, p in model.named_parameters()}
    for step in range(3):
        with torch.no_grad():
            for p in model.parameters():
                p.add_(torch.randn_like(p) * 0.01)
        ema.update()
        for n, p in model.named_parameters():
            ref[n] = (1.0 - 0.999) * p.data + 0.999 * ref[n]
    torch.cuda.synchronize()
    for n, p in ema_model.named_parameters():
        assert p.is_cuda and torch.equal(p.data, ref[n]), n
        assert ema.shadow[n].data_ptr() == p.data_ptr()
    keep = {n: p.data for n, p in model.named_parameters()}
    ema.apply_shadow()
    assert all(p.data.data_ptr() == ema.shadow[n].data_ptr() for n, p in model.named_parameters())
    ema.restore()
    assert all(p.data.data_ptr() == keep[n].data_ptr() for n, p in model.named_parameters())


def test_softmatch_mask_data_parallel_view():
    """C4 for SoftMatch (dist_align.py:40-42, srsoftmatch/utils.py:33-34): DistAlign's EMA and the mean / variance EMA
    integrate the rows of ALL ranks, the weights are formed for the local rows.  Emulated in one process for 'rank 1' of 2
    through the hook's `gathered` testing hook, against the reference state machines fed the same way."""
    from oracle import ssl_oracle as O
    from semireward_b200 import detgen
    from semireward_b200.core.hooks import DistAlignEMAHook, SoftMatchWeightingHook
    B, C, W, rank = 8, 100, 2, 1
    st, da = O.SoftMatchState(C, 2, 0.9), O.DistAlignState(C, 0.9)
    hook, dah = SoftMatchWeightingHook(C, 2, 0.9, device="cuda"), DistAlignEMAHook(C, 0.9, device="cuda")
    alg = _Alg()
    alg.hooks_dict["DistAlignHook"] = dah
    local = slice(rank * B, (rank + 1) * B)
    for call in range(5):
        logits_all = torch.from_numpy(detgen.normal("sm_dp_logits", (W * B, C), 110 + call)) * (1.0 + 0.5 * call)
        probs_all = torch.softmax(logits_all, dim=-1)
        align = call % 2 == 0
        used_all = da.dist_align(probs_all) if align else probs_all      # every rank aligns with the same (gathered) p_model
        maxp_all = used_all.max(dim=-1)[0]
        # reference update() on the gathered rows, weights for the local rows (srsoftmatch/utils.py:31-77)
        mu, var = torch.mean(maxp_all).item(), torch.var(maxp_all, unbiased=True).item()
        st.prob_max_mu_t = st.m * st.prob_max_mu_t + (1 - st.m) * mu
        st.prob_max_var_t = st.m * st.prob_max_var_t + (1 - st.m) * var
        ref_w = torch.exp(-((torch.clamp(maxp_all[local] - st.prob_max_mu_t, max=0.0) ** 2) / (2 * st.prob_max_var_t / (st.n_sigma ** 2))))

        def gathered(kind, t):
            full = (probs_all if kind == "probs" else maxp_all).cuda().clone()
            full[local] = t     # the local rows come from the kernel itself
            return full
        w = hook.masking(alg, logits_all[local].cuda(), softmax_x_ulb=True, dist_align=align, pseudo_from_probs=not align, gathered=gathered)
        torch.cuda.synchronize()
        assert abs(hook.prob_max_mu_t.item() - float(st.prob_max_mu_t)) < 1e-6
        assert abs(hook.prob_max_var_t.item() - float(st.prob_max_var_t)) < 1e-6
        if align:
            assert (dah.p_model.cpu() - da.p_model).abs().max().item() < 1e-6
        assert (w.cpu() - ref_w).abs().max().item() < 2e-6, f"call {call}"


@pytest.mark.parametrize("B,C,D", [(8, 100, 384), (24, 10, 768), (128, 1000, 768)])
def test_fused_stage2_epilogue_equals_separate_launches(B, C, D):
    """north_star's "one fused epilogue kernel": srw_ssl_loss with the Rewarder inside (forward -> mean-threshold mask2 -> masked
    consistency CE -> dlogits, one launch) against srw_rewarder_fwd followed by srw_ssl_loss — bit-identical (same device code),
    and against the oracle's functions.  B = 8 / 24 keep the Rewarder's intermediates in shared memory, B = 128 in the workspace."""
    from oracle import ssl_oracle as O
    from semireward_b200 import detgen
    from semireward_b200.algorithms.semireward import Rewarder, label_dim
    from semireward_b200.algorithms.srflexmatch import _ssl_loss_native
    R = Rewarder(label_dim(C), 128, D)
    rp = {n: torch.from_numpy(detgen.fill_param("rewarder." + n, s, 0)) for n, s in O.rewarder_param_shapes(D, C)}
    with torch.no_grad():
        for n, q in R.named_parameters():
            q.copy_(rp[n])
    R = R.cuda()
    feats = torch.from_numpy(detgen.normal("f", (B, D), 3)).cuda()
    llb = torch.from_numpy(detgen.normal("llb", (8, C), 4, std=2.0)).cuda()
    ls = torch.from_numpy(detgen.normal("ls", (B, C), 5, std=2.0)).cuda()
    y = torch.from_numpy(detgen.integers("y", (8,), 0, C, 6)).cuda()
    pseudo = torch.from_numpy(detgen.integers("p", (B,), 0, C, 7)).cuda()
    mask = (torch.from_numpy(detgen.normal("m", (B,), 8)) > -0.3).float().cuda()
    dl1, ds1, dl2, ds2 = (torch.empty(8, C, device="cuda"), torch.empty(B, C, device="cuda"), torch.empty(8, C, device="cuda"), torch.empty(B, C, device="cuda"))
    reward = R(feats, pseudo)
    lo_a, m2_a = _ssl_loss_native(llb, ls, y, pseudo, mask, reward.view(-1), 1.0, dl1, ds1)
    lo_b, m2_b = _ssl_loss_native(llb, ls, y, pseudo, mask, None, 1.0, dl2, ds2, rewarder=R, feats=feats)
    torch.cuda.synchronize()
    assert torch.equal(m2_a, m2_b) and torch.equal(lo_a[:4], lo_b[:4]) and torch.equal(dl1, dl2) and torch.equal(ds1, ds2)
    r_ref = O.rewarder_forward(rp, feats.cpu(), pseudo.cpu())
    assert (reward.cpu() - r_ref).abs().max().item() < 2e-6
    m2_ref = torch.where(r_ref >= r_ref.mean(), 1, 0).squeeze().float()
    tied = ((r_ref - r_ref.mean()).abs() <= 4e-7 * r_ref.mean().abs()).flatten()
    assert torch.equal(m2_b.cpu()[~tied], m2_ref[~tied])
    if not tied.any():
        unsup_ref = O.consistency_loss(ls.cpu(), pseudo.cpu(), mask.cpu(), m2_ref)
        assert abs(float(lo_b[1]) - float(unsup_ref)) < 1e-5 * max(1.0, abs(float(unsup_ref)))
