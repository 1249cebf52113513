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
