"""-m gpu: the audio path (SURVEY.md §8a row a4 + §8f rank 3, BASELINE configs[4]) — native ClassificationHubert (srw_hubert_forward /
srw_hubert_backward: conv stem as GEMMs over overlapping views, GroupNorm, weight-normalised grouped positional conv, post-LN encoder)
and the SSL step with `use_cat: False` against oracle/hubert_oracle.py, which is pinned to the live reference + Hugging Face HubertModel
(tests/test_hubert_oracle.py).  Gates: logits / feat / losses 1e-3, pseudo-labels and masks bit-exact, gradients 1e-3 relative.
Stochastic passes are compared with the SAME decisions on both sides: counter-based dropout bits, injected LayerDrop skips and
SpecAugment frames."""
import functools

import numpy as np
import pytest
import torch

from golden_cases import hubert_small_cfg

pytestmark = pytest.mark.gpu


def _step_cfg(cfg):
    from oracle import ssl_oracle as O
    return O.StepConfig(algorithm=cfg["algorithm"], num_classes=cfg["num_classes"], ulb_dest_len=cfg["ulb_dest_len"], p_cutoff=cfg["p_cutoff"],
                        thresh_warmup=cfg["thresh_warmup"], start_timing=cfg["start_timing"], N_k=cfg["N_k"], num_train_iter=cfg["num_train_iter"],
                        num_warmup_iter=cfg["num_warmup_iter"], lr=cfg["lr"], weight_decay=cfg["weight_decay"], layer_decay=cfg["layer_decay"],
                        sr_lr=cfg["sr_lr"], feature_dim=cfg["feature_dim"])


def _audio_batch(cfg, it, samples, seed=1):
    from semireward_b200 import detgen
    b = detgen.audio_batch(cfg["batch_size"], cfg["uratio"], cfg["num_classes"], cfg["ulb_dest_len"], samples=samples, seed=seed, step=it)
    return {k: torch.from_numpy(v) for k, v in b.items()}


def _build_native(cfg, layers, head_gain, drop=0.0, seed=0):
    import semireward_b200 as S
    from semireward_b200 import detgen
    args = S.get_config(dict(cfg, gpu=0))
    builder = functools.partial(S.get_net_builder("hubert_base"), num_hidden_layers=layers, feat_proj_dropout=drop, hidden_dropout=drop,
                                attention_dropout=drop, activation_dropout=drop, pooled_dropout=drop, layerdrop=0.0, apply_spec_augment=False)
    alg = S.get_algorithm(args, builder, None, None)
    with torch.no_grad():
        for prefix, mod in (("", alg.model), ("rewarder.", alg.rewarder), ("generator.", alg.generator)):
            for n, p in mod.named_parameters():
                p.copy_(torch.from_numpy(detgen.fill_param(prefix + n, p.shape, seed)))
                if prefix == "" and n == "classifier.2.weight":
                    p.mul_(head_gain)
    alg.model = alg.model.cuda(0).train()
    alg.rewarder, alg.generator = alg.rewarder.cuda(0), alg.generator.cuda(0)
    return alg


def test_hubert_state_dict_keys_match_the_oracle_table():
    from oracle import hubert_oracle as HO
    import semireward_b200 as S
    net = S.get_net_builder("hubert_base")(num_classes=10, num_hidden_layers=2)
    hc = HO.HubertCfg(layers=2, num_classes=10)
    assert [(n, tuple(p.shape)) for n, p in net.named_parameters()] == hc.param_shapes()
    assert list(net.state_dict()) == [n for n, _ in hc.param_shapes()]


@pytest.mark.parametrize("layers,samples,S,drop,stoch", [(2, 8000, 3, 0.0, False), (2, 16000, 2, 0.1, True), (1, 64000, 2, 0.0, False), (3, 8000, 4, 0.0, True)])
def test_hubert_backbone_forward_backward_vs_oracle(layers, samples, S, drop, stoch):
    """ClassificationHubert forward + backward on one batch.  stoch: SpecAugment frames and a LayerDrop skip injected on both sides
    (two model calls in one launch, the second skips layer 0); drop > 0: the counter masks of stream key call_key(seed, call)."""
    from oracle import hubert_oracle as HO
    from oracle.bert_oracle import call_key as okey
    from semireward_b200 import detgen
    from semireward_b200.nets.bert import call_key
    cfg = hubert_small_cfg(num_classes=10)
    hc = HO.HubertCfg(layers=layers, num_classes=10, feat_proj_dropout=drop, hidden_dropout=drop, attention_dropout=drop, activation_dropout=drop,
                      pooled_dropout=drop)
    alg = _build_native(cfg, layers, 1.0, drop=drop)
    net = alg.model
    Fr = hc.frames(samples)
    assert net.frames(samples) == Fr
    p = {n: torch.from_numpy(detgen.fill_param(n, s, 0)).requires_grad_(True) for n, s in hc.param_shapes()}
    x = torch.from_numpy(detgen.normal("clip", (S, samples), 9))
    assert okey(7, 0) == call_key(7, 0)
    # the launch = two model calls: clips [0, S1) and [S1, S)
    S1 = S // 2 if stoch else S
    mask = None
    skips = [(), ()]
    if stoch:
        mask = torch.zeros(S, Fr, dtype=torch.bool)
        mask[0, 2:7] = True
        mask[S - 1, Fr - 3:] = True
        skips = [(), (0,)]
    lg_ref, ft_ref = [], []
    for c, (a, b) in enumerate(((0, S1), (S1, S))):
        if b <= a:
            continue
        l_, f_ = HO.hubert_forward(p, x[a:b], hc, mask_time_indices=None if mask is None else mask[a:b], skip_layers=skips[c],
                                   drop_key=okey(7, c) if drop > 0 else None)
        lg_ref.append(l_); ft_ref.append(f_)
    logits_ref, feat_ref = torch.cat(lg_ref), torch.cat(ft_ref)
    g = torch.Generator().manual_seed(3)
    dlog = torch.randn(S, 10, generator=g)
    dft = torch.randn(S, 768, generator=g) * 0.01
    (logits_ref * dlog).sum().add((feat_ref * dft).sum()).backward()
    dev = torch.device("cuda")
    inp = net.concat_inputs([x.cuda()], dev)
    spec = None
    if stoch or drop > 0:
        keys = [call_key(7, 0)] * S1 + [call_key(7, 1)] * (S - S1)
        rows = list(range(S1)) + list(range(S - S1))
        sk = np.zeros((2, layers), dtype=np.uint8)
        for c in range(2):
            for l in skips[c]:
                sk[c, l] = 1
        spec = dict(keys=torch.from_numpy(np.asarray(keys, dtype=np.uint32).view(np.int32).copy()) if drop > 0 else None,
                    rows=torch.tensor(rows, dtype=torch.int32) if drop > 0 else None,
                    segments=np.asarray([0, S1, S], dtype=np.int32), skip=sk,
                    masks=[None if mask is None else mask[:S1].numpy(), None if mask is None else mask[S1:].numpy()])
    lg, ft, h = net.forward_native(inp, grad_batch=S, drop_scale=spec)
    torch.cuda.synchronize()
    e_l, e_f = (lg.cpu() - logits_ref.detach()).abs().max().item(), (ft.cpu() - feat_ref.detach()).abs().max().item()
    print(f"hubert layers {layers} samples {samples} ({Fr} frames) drop {drop} stoch {stoch}: logits err {e_l:.2e} feat err {e_f:.2e}")
    assert e_l < 1e-3 and e_f < 1e-3
    flat, views = net.backward_native(h, dlog.cuda(), dfeat=dft.cuda())
    torch.cuda.synchronize()
    names = {id(q): n for n, q in net.named_parameters()}
    worst, worst_n, rows_ = 0.0, "", []
    gmax = max(q.grad.abs().max().item() for q in p.values() if q.grad is not None)
    for q, v in zip(net._grad_params(), views):
        n = names[id(q)]
        gr = p[n].grad
        if gr is None:   # unused parameter (masked_spec_embed without SpecAugment): the native gradient is exactly zero
            assert v.abs().max().item() == 0.0, n
            continue
        sc = gr.abs().max().item()
        err = (v.cpu() - gr).abs().max().item() / max(sc, 1e-12)
        if sc < 1e-9 or n.endswith("attention.k_proj.bias"):
            # mathematically zero gradient (softmax is invariant to a per-query shift of the scores): both sides hold rounding noise
            assert v.abs().max().item() < 1e-5 and sc < 1e-5, (n, v.abs().max().item(), sc)
            continue
        rows_.append((err, n, sc))
        if err > worst:
            worst, worst_n = err, n
    rows_.sort(reverse=True)
    print(f"   worst gradient rel err {worst:.2e} ({worst_n}); largest |grad| {gmax:.2e}")
    for e, n, sc in rows_[:6]:
        print(f"      {n}: rel err {e:.2e} max |grad| {sc:.2e}")
    for e, n, sc in rows_:   # tensors whose gradient is itself cancellation noise are held to 1e-3 of the largest gradient instead
        assert e < 1e-3 or e * sc < 1e-5 * gmax, (n, e, sc, gmax)


@pytest.mark.parametrize("algorithm,over,hg,drop", [("srflexmatch", dict(), 4.0, 0.0), ("srfixmatch", dict(p_cutoff=0.5), 4.0, 0.0),
                                                    ("srflexmatch", dict(), 4.0, 0.1)])
def test_hubert_ssl_steps_vs_oracle(algorithm, over, hg, drop):
    """SRFlexMatch (the configs[4] algorithm) / SRFixMatch steps with `use_cat: False` on the audio backbone: stage 1, the gap step,
    stage 2 with and without an SR update.  drop = 0.1: every dropout site on, the same counter streams on both sides (three calls
    per pass, K passes in stage 2: the native batched route against the oracle's sequential passes)."""
    from oracle import hubert_oracle as HO
    from test_train_step_gpu import _grad_tap, _mask2_tie
    cfg = hubert_small_cfg(algorithm=algorithm, num_train_iter=16, start_timing=2, **over)
    hc = HO.HubertCfg(layers=2, num_classes=cfg["num_classes"], feat_proj_dropout=drop, hidden_dropout=drop, attention_dropout=drop,
                      activation_dropout=drop, pooled_dropout=drop)
    orc = HO.build_det_hubert_oracle(hc, _step_cfg(cfg), seed=0, head_gain=hg)
    alg = _build_native(cfg, 2, hg, drop=drop)
    if drop > 0:
        orc.drop_seed = 5
        alg.model.dropout_seed = 5
    tap = _grad_tap(alg)
    for it in range(6):
        batch = _audio_batch(cfg, it, 8000)
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
        tie = _mask2_tie(rec)
        if not tie:
            if "dg_mask2" in rec:
                assert torch.equal(alg._last_mask2.cpu(), rec["dg_mask2"]), f"it {it}: mask2 differs"
            for kn, ko in (("train/unsup_loss", "unsup_loss"), ("train/total_loss", "total_loss")):
                assert abs(ld[kn] - float(rec[ko])) < 1e-3, f"it {it} {ko}: {ld[kn]} vs {float(rec[ko])}"
            worst, wn = 0.0, ""
            gmax = max(g.abs().max().item() for g in ref_grads.values() if g is not None)
            for n, q in alg.model.named_parameters():
                gr = ref_grads[n]
                if gr is None:
                    continue
                sc = gr.abs().max().item()
                if sc < 1e-9 or n.endswith("attention.k_proj.bias"):
                    continue
                e = (tap[n].cpu() - gr).abs().max().item() / sc
                if e > worst and not e * sc < 1e-5 * gmax:
                    worst, wn = e, n
            print(f"hubert {algorithm} drop {drop} it {it}: total {ld['train/total_loss']:.5f} (oracle {float(rec['total_loss']):.5f}) K {rec.get('K', 0)} "
                  f"grad rel err {worst:.2e} ({wn})")
            assert worst < 1e-3, f"it {it}: gradient error {worst} ({wn})"
        else:
            print(f"hubert {algorithm} it {it}: rewards tie with their mean (mask2-dependent checks skipped)")
        with torch.no_grad():   # resync from the oracle
            for n, q in alg.model.named_parameters():
                q.copy_(orc.p[n].detach())
            for n, q in alg.rewarder.named_parameters():
                q.copy_(orc.rp[n].detach())
        alg.model.mark_weights_updated()


def test_hubert_ssl_steps_with_layerdrop_and_specaugment_vs_oracle():
    """The two structural sources of randomness of HubertModel's train mode at the level of the SSL step: every model call of the run
    gets its own LayerDrop pattern and SpecAugment frames (injected identically into the oracle's calls and into the native launch's
    call segments), dropout 0.1 on top.  Stage 1, the gap step and stage 2 (K + 1 passes = 3 (K + 1) calls in ONE native launch whose
    segments skip different layers, against the oracle's sequential calls)."""
    from oracle import hubert_oracle as HO
    from test_train_step_gpu import _grad_tap, _mask2_tie
    drop, layers, samples = 0.1, 3, 8000
    cfg = hubert_small_cfg(algorithm="srflexmatch", num_train_iter=16, start_timing=2)
    hc = HO.HubertCfg(layers=layers, num_classes=cfg["num_classes"], feat_proj_dropout=drop, hidden_dropout=drop, attention_dropout=drop,
                      activation_dropout=drop, pooled_dropout=drop)
    orc = HO.build_det_hubert_oracle(hc, _step_cfg(cfg), seed=0, head_gain=4.0)
    alg = _build_native(cfg, layers, 4.0, drop=drop)
    net = alg.model
    orc.drop_seed = 5
    net.dropout_seed = 5
    Fr = hc.frames(samples)
    nl, nu = cfg["batch_size"], cfg["batch_size"] * cfg["uratio"]
    rng = np.random.default_rng(7)

    def decisions(call):   # deterministic per call index: ~1 layer in 3 skipped, two short spans per clip
        r = np.random.default_rng(1000 + call)
        skip = (r.random(layers) < 0.34).astype(np.uint8)
        n = nl if call % 3 == 0 else nu
        mask = np.zeros((n, Fr), dtype=bool)
        for b in range(n):
            for s0 in r.choice(Fr - 3, 2, replace=False):
                mask[b, s0:s0 + 3] = True
        return skip, mask

    orc.stochastic_inputs = lambda call: (torch.from_numpy(decisions(call)[1]), tuple(int(i) for i in np.nonzero(decisions(call)[0])[0]))
    tap = _grad_tap(alg)
    seen_skip = 0
    for it in range(5):
        K = 0 if it <= cfg["start_timing"] else int(max(8, 1 + cfg["num_train_iter"] / it))
        for c in range(net._calls, net._calls + 3 * (K + 1)):
            sk, mk = decisions(c)
            net.set_call_draws(c, layer_skip=sk, mask_time=mk)
            seen_skip += int(sk.sum())
        batch = _audio_batch(cfg, it, samples)
        rec = orc.train_step(dict(batch), it)
        ref_grads = orc.param_update()
        alg.it = it
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**batch))
        alg.call_hook("after_train_step")
        torch.cuda.synchronize()
        assert net._calls == orc.calls, (it, net._calls, orc.calls)
        ld = alg.log_dict
        assert abs(ld["train/sup_loss"] - float(rec["sup_loss"])) < 1e-3, (it, ld, float(rec["sup_loss"]))
        assert torch.equal(alg._last_pseudo_label.cpu(), rec["pseudo"]) and torch.equal(alg._last_mask.cpu(), rec["mask"]), it
        if not _mask2_tie(rec):
            for kn, ko in (("train/unsup_loss", "unsup_loss"), ("train/total_loss", "total_loss")):
                assert abs(ld[kn] - float(rec[ko])) < 1e-3, f"it {it} {ko}: {ld[kn]} vs {float(rec[ko])}"
            worst, wn = 0.0, ""
            gmax = max(g.abs().max().item() for g in ref_grads.values() if g is not None)
            for n, q in net.named_parameters():
                gr = ref_grads[n]
                if gr is None:
                    continue
                sc = gr.abs().max().item()
                if sc < 1e-9 or n.endswith("attention.k_proj.bias"):
                    continue
                e = (tap[n].cpu() - gr).abs().max().item() / sc
                if e > worst and not e * sc < 1e-5 * gmax:
                    worst, wn = e, n
            print(f"hubert layerdrop+specaugment it {it}: total {ld['train/total_loss']:.5f} (oracle {float(rec['total_loss']):.5f}) K {rec.get('K', 0)} grad rel err {worst:.2e} ({wn})")
            assert worst < 1e-3, (it, worst, wn)
            assert tap["model.masked_spec_embed"].abs().max().item() > 0      # SpecAugment frames feed their gradient to the embedding
        with torch.no_grad():
            for n, q in net.named_parameters():
                q.copy_(orc.p[n].detach())
            for n, q in alg.rewarder.named_parameters():
                q.copy_(orc.rp[n].detach())
        net.mark_weights_updated()
    assert seen_skip > 0
