"""-m gpu: the convolutional path (SURVEY.md §8a row a5 + §8f rank 3, BASELINE configs[0]) — native WideResNet (srw_wrn_forward /
srw_wrn_backward: 3x3 convolutions as K-segmented GEMMs over zero-bordered NHWC planes, train-mode BatchNorm over all rows of the
launch, running statistics, fused SGD-nesterov) and the SSL step with `use_cat: True` against oracle/wrn_oracle.py, which is pinned
bit for bit to the live reference (tests/test_wrn_oracle.py).  Gates: logits / feat / losses 1e-3, pseudo-labels and masks bit-exact,
gradients 1e-3 relative, BatchNorm buffers 1e-5."""
import functools

import pytest
import torch

from golden_cases import wrn_small_cfg

pytestmark = pytest.mark.gpu


def _step_cfg(cfg):
    from oracle import ssl_oracle as O
    return O.StepConfig(algorithm=cfg["algorithm"], num_classes=cfg["num_classes"], ulb_dest_len=cfg["ulb_dest_len"], p_cutoff=cfg["p_cutoff"],
                        thresh_warmup=cfg["thresh_warmup"], start_timing=cfg["start_timing"], N_k=cfg["N_k"], num_train_iter=cfg["num_train_iter"],
                        num_warmup_iter=cfg["num_warmup_iter"], lr=cfg["lr"], weight_decay=cfg["weight_decay"], layer_decay=cfg["layer_decay"],
                        sr_lr=cfg["sr_lr"], feature_dim=cfg["feature_dim"])


def _batch(cfg, it, img=32):
    from semireward_b200 import detgen
    b = detgen.ssl_batch(cfg["batch_size"], cfg["uratio"], cfg["num_classes"], cfg["ulb_dest_len"], img_size=img, seed=1, step=it)
    return {k: torch.from_numpy(v) for k, v in b.items()}


def _build_native(cfg, depth, head_gain, widen=2, img=32, seed=0, slope=0.1):
    import semireward_b200 as S
    from semireward_b200 import detgen
    args = S.get_config(dict(cfg, gpu=0))
    from semireward_b200.nets.wrn import WideResNet
    assert S.get_net_builder("wrn_28_2").__module__ == WideResNet.__module__
    builder = functools.partial(WideResNet, first_stride=1, depth=depth, widen_factor=widen, img_size=img, leaky_slope=slope)   # wrn_28_2 with another depth / width
    alg = S.get_algorithm(args, builder, None, None)
    with torch.no_grad():
        for prefix, mod in (("", alg.model), ("rewarder.", alg.rewarder), ("generator.", alg.generator)):
            for n, p in mod.named_parameters():
                p.copy_(torch.from_numpy(detgen.fill_param(prefix + n, p.shape, seed)))
                if prefix == "" and n == "classifier.weight":
                    p.mul_(head_gain)
    alg.model = alg.model.cuda(0).train()
    alg.rewarder, alg.generator = alg.rewarder.cuda(0), alg.generator.cuda(0)
    return alg


def test_wrn_state_dict_keys_match_the_reference_tree():
    from oracle import wrn_oracle as WO
    import semireward_b200 as S
    net = S.get_net_builder("wrn_28_2")(num_classes=100)
    wc = WO.WRNCfg(num_classes=100)
    assert [(n, tuple(p.shape)) for n, p in net.named_parameters()] == wc.param_shapes()
    bufs = [n for n, _ in net.named_buffers()]
    want = []
    for name, _ in wc.bn_names():
        want += [name + ".running_mean", name + ".running_var", name + ".num_batches_tracked"]
    assert sorted(bufs) == sorted(want) and len(net.state_dict()) == 81 + 75
    assert sorted(net.no_weight_decay()) == sorted(n for n, _ in wc.param_shapes() if "bn" in n or "bias" in n)


def _grad_gate(rows_, views_by_name, p, gmax, strict):
    """strict: every tensor within 1e-3 of its own maximum (or below 1e-5 of the step's largest gradient: mathematically-zero tensors
    such as the stem bias in front of a BatchNorm).  Otherwise the LeakyReLU-kink gate: an implementation whose pre-activations differ
    by 1e-5 relative (bf16x3 products) takes the other slope on the few elements with |u| < 1e-5 |u|_typ; ONE such element of a layer
    with R rows moves the bias-gradient entry of its channel, and that channel's slice of the preceding convolution's weight gradient,
    by ~1 / sqrt(R) of their size (3 % at R = 1200), and everything upstream of it a little — in any two fp32 implementations.  So the
    gate is the relative L2 error per tensor (<= 3e-2; a wrong slope factor or a wrong operand gives O(1)) and a bound on the worst
    entry; the kink-free runs (slope 1.0) hold the same kernels to the strict gate."""
    for e, n, sc in rows_:
        if e * sc < 1e-5 * gmax:
            continue
        if strict:
            assert e < 1e-3, (n, e, sc, gmax)
            continue
        v, gr = views_by_name[n], p[n].grad
        l2 = ((v - gr).norm() / gr.norm().clamp_min(1e-20)).item()
        assert l2 < 3e-2 and e < 0.15, (n, e, l2)


@pytest.mark.parametrize("depth,widen,S,img,grad_rows,slope", [(10, 2, 6, 32, 6, 0.1), (16, 2, 5, 32, 3, 0.1), (28, 2, 12, 32, 8, 0.1), (10, 4, 3, 16, 3, 0.1),
                                                               (28, 2, 12, 32, 8, 1.0), (16, 4, 4, 16, 2, 1.0)])
def test_wrn_backbone_forward_backward_vs_oracle(depth, widen, S, img, grad_rows, slope):
    """WideResNet forward + backward on one batch in train mode: logits, features, every gradient, and the BatchNorm buffers after the
    call.  grad_rows < S: the tail rows carry no logit gradient but take part in every BatchNorm backward (the weak rows of a step).
    slope = 1.0 removes the LeakyReLU kink on both sides: every kernel of the backward is then held to the strict 1e-3 gate at full
    depth; with the reference's slope 0.1 the gradients go through the kink gate (_grad_gate)."""
    from oracle import wrn_oracle as WO
    from semireward_b200 import detgen
    cfg = wrn_small_cfg(num_classes=100)
    wc = WO.WRNCfg(depth=depth, widen=widen, num_classes=100, slope=slope)
    alg = _build_native(cfg, depth, 1.0, widen=widen, img=img, slope=slope)
    net = alg.model
    p = {n: torch.from_numpy(detgen.fill_param(n, s, 0)).requires_grad_(True) for n, s in wc.param_shapes()}
    buf = WO.new_bn_buffers(wc)
    x = torch.from_numpy(detgen.normal("imgs", (S, 3, img, img), 5))
    logits_ref, feat_ref = WO.wrn_forward(p, buf, x, wc, training=True)
    g = torch.Generator().manual_seed(3)
    dlog = torch.randn(grad_rows, 100, generator=g)
    dft = torch.randn(grad_rows, 64 * widen, generator=g) * 0.01
    (logits_ref[:grad_rows] * dlog).sum().add((feat_ref[:grad_rows] * dft).sum()).backward()
    dev = torch.device("cuda")
    lg, ft, h = net.forward_native(net.concat_inputs([x.cuda()], dev), grad_batch=grad_rows)
    torch.cuda.synchronize()
    e_l, e_f = (lg.cpu() - logits_ref.detach()).abs().max().item(), (ft.cpu() - feat_ref.detach()).abs().max().item()
    print(f"wrn-{depth}-{widen} S {S} img {img} slope {slope}: logits err {e_l:.2e} feat err {e_f:.2e}")
    assert e_l < 1e-3 and e_f < 1e-3
    worst_b = 0.0
    for n, b in net.named_buffers():
        if n.endswith("num_batches_tracked"):
            assert int(b) == 1, n
            continue
        worst_b = max(worst_b, (b.cpu() - buf[n]).abs().max().item())
    assert worst_b < 1e-5, worst_b
    flat, views = net.backward_native(h, dlog.cuda(), dfeat=dft.cuda())
    torch.cuda.synchronize()
    names = {id(q): n for n, q in net.named_parameters()}
    rows_, by_name = [], {}
    gmax = max(q.grad.abs().max().item() for q in p.values() if q.grad is not None)
    for q, v in zip(net._grad_params(), views):
        n = names[id(q)]
        gr = p[n].grad
        assert gr is not None, n
        sc = gr.abs().max().item()
        by_name[n] = v.cpu()
        rows_.append(((by_name[n] - gr).abs().max().item() / max(sc, 1e-12), n, sc))
    rows_.sort(reverse=True)
    print(f"   worst gradient rel err {rows_[0][0]:.2e} ({rows_[0][1]}); largest |grad| {gmax:.2e}; bn buffers {worst_b:.1e}")
    for e, n, sc in rows_[:5]:
        print(f"      {n}: rel err {e:.2e} max |grad| {sc:.2e}")
    _grad_gate(rows_, by_name, p, gmax, strict=(slope == 1.0))
    dead = {n for n, q in net.named_parameters() if id(q) not in {id(t) for t in net._grad_params()}}
    assert dead == {"block2.layer.0.bn1.weight", "block2.layer.0.bn1.bias", "block3.layer.0.bn1.weight", "block3.layer.0.bn1.bias"}
    assert all(p[n].grad is None for n in dead)


def test_wrn_eval_mode_uses_running_statistics():
    from oracle import wrn_oracle as WO
    from semireward_b200 import detgen
    cfg = wrn_small_cfg(num_classes=100)
    wc = WO.WRNCfg(depth=10, num_classes=100)
    alg = _build_native(cfg, 10, 1.0)
    net = alg.model
    p = {n: torch.from_numpy(detgen.fill_param(n, s, 0)) for n, s in wc.param_shapes()}
    buf = WO.new_bn_buffers(wc)
    for n in buf:   # non-trivial running statistics on both sides
        buf[n] = torch.from_numpy(detgen.uniform(n, buf[n].shape, 3, 0.5, 1.5)) if n.endswith("var") else torch.from_numpy(detgen.normal(n, buf[n].shape, 3, 0.1))
    with torch.no_grad():
        for n, b in net.named_buffers():
            if n in buf:
                b.copy_(buf[n])
    x = torch.from_numpy(detgen.normal("imgs", (4, 3, 32, 32), 6))
    want_l, want_f = WO.wrn_forward(p, {k: v.clone() for k, v in buf.items()}, x, wc, training=False)
    net.eval()
    out = net(x.cuda())
    assert (out["logits"].cpu() - want_l).abs().max().item() < 1e-3 and (out["feat"].cpu() - want_f).abs().max().item() < 1e-3
    for n, b in net.named_buffers():
        if n in buf:
            assert torch.equal(b.cpu(), buf[n]), n      # eval never touches the buffers


@pytest.mark.parametrize("algorithm,over,hg", [("srflexmatch", dict(), 4.0), ("srfixmatch", dict(p_cutoff=0.2), 4.0)])
def test_wrn_ssl_steps_vs_oracle(algorithm, over, hg):
    """SRFlexMatch (BASELINE configs[0]) / SRFixMatch steps on WRN with `use_cat: True` and SGD-nesterov: stage 1, the gap step, stage 2
    with and without an SR update; parameters after the fused SGD step and the BatchNorm buffers (which advance 1 + K times per
    stage-2 step) against the oracle."""
    from oracle import wrn_oracle as WO
    from test_train_step_gpu import _grad_tap, _mask2_tie
    cfg = wrn_small_cfg(algorithm=algorithm, num_train_iter=16, start_timing=2, **over)
    wc = WO.WRNCfg(depth=10, num_classes=cfg["num_classes"])
    orc = WO.build_det_wrn_oracle(wc, _step_cfg(cfg), seed=0, head_gain=hg)
    alg = _build_native(cfg, 10, hg)
    tap = _grad_tap(alg)
    for it in range(6):
        batch = _batch(cfg, it)
        rec = orc.train_step(dict(batch), it)
        ref_grads = orc.param_update()
        alg.it = it
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**batch))
        alg.call_hook("after_train_step")
        torch.cuda.synchronize()
        ld = alg.log_dict
        assert abs(ld["train/sup_loss"] - float(rec["sup_loss"])) < 1e-3, (it, ld, float(rec["sup_loss"]))
        assert torch.equal(alg._last_pseudo_label.cpu(), rec["pseudo"]), f"it {it}: pseudo labels differ"
        assert torch.equal(alg._last_mask.cpu(), rec["mask"]), f"it {it}: mask differs"
        worst_b = 0.0
        for n, b in alg.model.named_buffers():
            if not n.endswith("num_batches_tracked"):
                worst_b = max(worst_b, (b.cpu() - orc.buf[n]).abs().max().item())
        assert worst_b < 1e-5, (it, worst_b)
        tie = _mask2_tie(rec)
        if not tie:
            if "dg_mask2" in rec:
                assert torch.equal(alg._last_mask2.cpu(), rec["dg_mask2"]), f"it {it}: mask2 differs"
            for kn, ko in (("train/unsup_loss", "unsup_loss"), ("train/total_loss", "total_loss")):
                assert abs(ld[kn] - float(rec[ko])) < 1e-3, f"it {it} {ko}: {ld[kn]} vs {float(rec[ko])}"
            worst, wn, worst_l2 = 0.0, "", 0.0
            gmax = max(g.abs().max().item() for g in ref_grads.values() if g is not None)
            for n, q in alg.model.named_parameters():
                gr = ref_grads[n]
                if gr is None:
                    assert n not in tap
                    continue
                sc = gr.abs().max().item()
                if sc < 1e-5 * gmax:
                    continue
                e = (tap[n].cpu() - gr).abs().max().item() / sc
                worst_l2 = max(worst_l2, ((tap[n].cpu() - gr).norm() / gr.norm()).item())
                if e > worst:
                    worst, wn = e, n
            # parameters after the fused SGD-nesterov step (momentum buffers included through the next steps)
            perr = max((q.detach().cpu() - orc.p[n].detach()).abs().max().item() for n, q in alg.model.named_parameters())
            print(f"wrn {algorithm} it {it}: total {ld['train/total_loss']:.5f} (oracle {float(rec['total_loss']):.5f}) K {rec.get('K', 0)} grad max err {worst:.2e} ({wn}) "
                  f"L2 {worst_l2:.2e} param err after SGD {perr:.2e} bn buffers {worst_b:.1e}")
            # LeakyReLU kink flips (see _grad_gate): single elements of a 20-image batch put ~1 / sqrt(rows) noise into a few entries
            assert worst_l2 < 2e-2 and worst < 0.1, f"it {it}: gradient error {worst} / L2 {worst_l2} ({wn})"
            assert perr < 1e-3, (it, perr)
        else:
            print(f"wrn {algorithm} it {it}: rewards tie with their mean (mask2-dependent checks skipped)")
        with torch.no_grad():   # resync parameters from the oracle (the momentum buffers stay the native ones)
            for n, q in alg.model.named_parameters():
                q.copy_(orc.p[n].detach())
            for n, q in alg.rewarder.named_parameters():
                q.copy_(orc.rp[n].detach())
        alg.model.mark_weights_updated()
