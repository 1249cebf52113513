"""CPU: the audio-path oracle (oracle/hubert_oracle.py: ClassificationHubert + the SSL step with use_cat False, SURVEY.md §8a
row a4 / BASELINE configs[4]) against golden vectors generated from the LIVE reference by tests/golden/make_golden_hubert.py
and — when /root/reference is mounted (build container) — bit for bit against the live reference and Hugging Face
HubertModel themselves.  The reference ships no fixtures for this path; these files are the pin.
Same torch CPU primitives in the same order on both sides, so the bar is float32 round-off; masks and integers bit-exact."""
import os

import numpy as np
import pytest
import torch

from golden_cases import HUBERT_CASES, HUBERT_SMALL, STEPS, hubert_small_cfg

GOLD = os.path.join(os.path.dirname(__file__), "golden")
LIVE = os.path.isdir("/root/reference/semilearn")


def _step_cfg(cfg):
    from oracle import ssl_oracle as O
    return O.StepConfig(algorithm=cfg["algorithm"], num_classes=cfg["num_classes"], ulb_dest_len=cfg["ulb_dest_len"], p_cutoff=cfg["p_cutoff"],
                        thresh_warmup=cfg["thresh_warmup"], start_timing=cfg["start_timing"], N_k=cfg["N_k"], num_train_iter=cfg["num_train_iter"],
                        num_warmup_iter=cfg["num_warmup_iter"], lr=cfg["lr"], weight_decay=cfg["weight_decay"], layer_decay=cfg["layer_decay"],
                        sr_lr=cfg["sr_lr"], feature_dim=cfg["feature_dim"])


def _batch(cfg, it):
    from semireward_b200 import detgen
    b = detgen.audio_batch(cfg["batch_size"], cfg["uratio"], cfg["num_classes"], cfg["ulb_dest_len"], samples=HUBERT_SMALL["samples"], seed=1, step=it)
    return {k: torch.from_numpy(v) for k, v in b.items()}


@pytest.mark.parametrize("name", sorted(HUBERT_CASES))
def test_hubert_oracle_matches_golden(name):
    from oracle import hubert_oracle as HO
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    spec = HUBERT_CASES[name]
    cfg = hubert_small_cfg(**spec["cfg"])
    orc = HO.build_det_hubert_oracle(HO.HubertCfg(layers=HUBERT_SMALL["layers"], num_classes=cfg["num_classes"]), _step_cfg(cfg), seed=0,
                                     head_gain=spec["head_gain"])
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    utils = []
    for it in range(STEPS):
        rec = orc.train_step(_batch(cfg, it), it)
        feat_lb = rec["feat_lb"].numpy().copy()
        orc.param_update()
        for key, val in (("loss", rec["total_loss"]), ("sup_loss", rec["sup_loss"]), ("unsup_loss", rec["unsup_loss"]), ("util_ratio", rec["util_ratio"])):
            np.testing.assert_allclose(np.float32(float(val)), gold[f"it{it}_{key}"], rtol=2e-5, atol=2e-6, err_msg=f"it{it} {key}")
        utils.append(float(rec["util_ratio"]))
        np.testing.assert_allclose(feat_lb, gold[f"it{it}_feat_lb"], rtol=0, atol=2e-4 if it else 2e-6, err_msg=f"it{it} feat_lb")
        probes = (("cls_bias", orc.p["classifier.2.bias"]), ("conv0_w0", orc.p["model.feature_extractor.conv_layers.0.conv.weight"][0, 0]),
                  ("pos_g", orc.p["model.encoder.pos_conv_embed.conv.parametrizations.weight.original0"].reshape(-1)),
                  ("q0_row0", orc.p["model.encoder.layers.0.attention.q_proj.weight"][0]))
        for key, val in probes:   # Adam turns noise-level gradient entries into +-lr moves: thread-count dependent reductions may flip a few
            np.testing.assert_allclose(val.detach().numpy(), gold[f"it{it}_{key}"], rtol=0, atol=1e-4, err_msg=f"it{it} {key}")
        rsum = sum(v.detach().double().sum().item() for v in orc.rp.values())
        assert abs(rsum - float(gold[f"it{it}_rewarder_sum"])) < 5e-2, (it, rsum)
        if cfg["algorithm"] == "srflexmatch":
            assert np.array_equal(orc.hook.selected_label.numpy(), gold[f"it{it}_selected_label"]), it
            assert np.array_equal(orc.hook.classwise_acc.numpy(), gold[f"it{it}_classwise_acc"]), it
    if name == "hubert_srfixmatch_l2":
        assert any(0.0 < u < 1.0 for u in utils), utils


def test_hubert_known_answers():
    from oracle import hubert_oracle as HO
    c = HO.HubertCfg()
    assert c.frames(64000) == 199 and HO.HubertCfg(layers=2).frames(4000) == 12                      # 199 frames for 4 s
    assert abs(c.fwd_flops_per_clip(64000) / 1e9 - 56.9) < 0.1
    shapes = c.param_shapes()
    assert sum(int(np.prod(s)) for _, s in shapes) == 94371712 + 768 * 768 + 768 + 10 * 768 + 10      # HubertModel(HubertConfig()) + classifier
    hp = HO.hubert_param_hparams(shapes, 12, 2e-5, 5e-4, 0.75)
    assert abs(hp["model.feature_extractor.conv_layers.3.conv.weight"][0] - 2e-5 * 0.75 ** 13) < 1e-20
    assert abs(hp["model.encoder.pos_conv_embed.conv.parametrizations.weight.original0"][0] - 2e-5 * 0.75 ** 13) < 1e-20
    assert hp["model.encoder.pos_conv_embed.conv.parametrizations.weight.original0"][1] == 5e-4     # the weight-norm gain is 3-D: decayed
    assert hp["model.masked_spec_embed"] == (2e-5, 0.0) and hp["model.encoder.layer_norm.weight"] == (2e-5, 0.0)
    assert abs(hp["model.encoder.layers.11.feed_forward.output_dense.weight"][0] - 2e-5 * 0.75) < 1e-20
    # masked_spec_embed is a parameter of the state dict that never receives a gradient in parity mode (SpecAugment off)
    from semireward_b200 import detgen
    c2 = HO.HubertCfg(layers=1)
    p = {n: torch.from_numpy(detgen.fill_param(n, s, 0)).requires_grad_(True) for n, s in c2.param_shapes()}
    logits, feat = HO.hubert_forward(p, torch.from_numpy(detgen.normal("clip", (2, 2000), 3)), c2)
    assert tuple(logits.shape) == (2, 10) and tuple(feat.shape) == (2, 768)
    logits.sum().backward()
    assert p["model.masked_spec_embed"].grad is None and p["model.feature_extractor.conv_layers.0.conv.weight"].grad is not None


@pytest.mark.skipif(not LIVE, reason="transformers pin / live reference only exist in the build container")
def test_hubert_forward_matches_huggingface_eager_and_sdpa():
    from transformers import HubertConfig, HubertModel
    from oracle import hubert_oracle as HO
    from semireward_b200 import detgen
    hc = HO.HubertCfg(layers=2)
    x = torch.from_numpy(detgen.normal("clip", (3, 4000), 5))
    for impl, tol in (("eager", 0.0), ("sdpa", 2e-6)):
        conf = HubertConfig(num_hidden_layers=2, hidden_dropout=0.0, activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, final_dropout=0.0,
                            layerdrop=0.0, apply_spec_augment=False, attn_implementation=impl)
        torch.manual_seed(0)
        m = HubertModel(conf).train()
        p = {"model." + n: v.detach().clone() for n, v in m.named_parameters()}
        assert list(p) == [n for n, _ in hc.param_shapes() if n.startswith("model.")]
        p.update({"classifier.0.weight": torch.zeros(768, 768), "classifier.0.bias": torch.zeros(768), "classifier.2.weight": torch.zeros(10, 768),
                  "classifier.2.bias": torch.zeros(10)})
        ref = m(x, return_dict=True)["last_hidden_state"].mean(1)
        _, feat = HO.hubert_forward(p, x, hc)
        err = (feat - ref).abs().max().item()
        assert err <= tol, (impl, err)


@pytest.mark.skipif(not LIVE, reason="transformers pin only exists in the build container")
def test_hubert_specaugment_and_layerdrop_injection_match_huggingface():
    """The structural randomness of a train-mode pass, injected explicitly on both sides: SpecAugment frames through
    `mask_time_indices` (HubertModel.forward takes them) and LayerDrop by pinning torch.rand([]) for the layer loop."""
    from transformers import HubertConfig, HubertModel
    from oracle import hubert_oracle as HO
    from semireward_b200 import detgen
    hc = HO.HubertCfg(layers=3)
    x = torch.from_numpy(detgen.normal("clip", (2, 4000), 9))
    conf = HubertConfig(num_hidden_layers=3, hidden_dropout=0.0, activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, final_dropout=0.0,
                        layerdrop=0.5, apply_spec_augment=True, attn_implementation="eager")   # explicit mask_time_indices: no random spans are drawn
    torch.manual_seed(0)
    m = HubertModel(conf).train()
    p = {"model." + n: v.detach().clone() for n, v in m.named_parameters()}
    p.update({"classifier.0.weight": torch.zeros(768, 768), "classifier.0.bias": torch.zeros(768), "classifier.2.weight": torch.zeros(10, 768),
              "classifier.2.bias": torch.zeros(10)})
    mask = torch.zeros(2, hc.frames(4000), dtype=torch.bool)
    mask[0, 2:5] = True
    mask[1, 7] = True
    draws = iter([0.9, 0.1, 0.7])                      # layer 1 is dropped (0.1 < layerdrop 0.5)
    orig = torch.rand
    torch.rand = lambda *a, **k: torch.tensor(next(draws)) if (len(a) == 1 and list(a[0]) == []) else orig(*a, **k)
    try:
        ref = m(x, mask_time_indices=mask, return_dict=True)["last_hidden_state"].mean(1)
    finally:
        torch.rand = orig
    _, feat = HO.hubert_forward(p, x, hc, mask_time_indices=mask, skip_layers=(1,))
    assert torch.equal(feat, ref), (feat - ref).abs().max().item()
    _, plain = HO.hubert_forward(p, x, hc)
    assert not torch.equal(plain, ref)


@pytest.mark.skipif(not LIVE, reason="live reference only exists in the build container")
def test_hubert_oracle_bit_exact_against_live_reference():
    import inspect
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from oracle import hubert_oracle as HO, ref_driver as R
    cfg = hubert_small_cfg(uratio=1)
    hf = dict(num_hidden_layers=2, hidden_dropout=0.0, activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, final_dropout=0.0, layerdrop=0.0,
              apply_spec_augment=False, attn_implementation="eager")
    alg = R.build_reference_algorithm(dict(cfg), net_kwargs=dict(hubert=hf, dropout=0.0))
    R.load_det_weights(alg, seed=0, head_gain=4.0)
    hc = HO.HubertCfg(layers=2, num_classes=cfg["num_classes"])
    assert [(n, tuple(p.shape)) for n, p in alg.model.named_parameters()] == hc.param_shapes()
    orc = HO.build_det_hubert_oracle(hc, _step_cfg(cfg), seed=0, head_gain=4.0)
    names = {id(p): n for n, p in alg.model.named_parameters()}
    for g in alg.optimizer.param_groups:
        for p in g["params"]:
            lr, wd = orc.hp[names[id(p)]]
            assert abs(g["lr"] - lr) < 1e-15 and g["weight_decay"] == wd, names[id(p)]
    for it in range(5):
        b = _batch(cfg, it)
        alg.it = it
        rb = {k: v for k, v in b.items() if k in inspect.signature(alg.train_step).parameters}
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**rb))
        alg.hooks_dict["ParamUpdateHook"].after_train_step(alg)
        rec = orc.train_step(b, it)
        orc.param_update()
        assert abs(alg.log_dict["train/total_loss"] - float(rec["total_loss"])) < 1e-6
        worst = max((p.detach() - orc.p[n].detach()).abs().max().item() for n, p in alg.model.named_parameters())
        assert worst < 1e-4, (it, worst)


def test_strided_conv_is_a_gemm_over_an_overlapping_view():
    """The layout plan for a native conv stem: with the signal stored time-major [T, C_in], row t of the im2col
    matrix is the k * C_in CONTIGUOUS elements starting at row s * t, so a strided Conv1d is one GEMM over an overlapping
    strided view (what a TMA tensor map with row stride s * C_in describes) against the weight reordered to [C_out, k, C_in]."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    for cin, cout, k, s, T in ((512, 512, 3, 2, 200), (1, 512, 10, 5, 4000), (512, 512, 2, 2, 99)):
        x = torch.randn(2, cin, T, generator=g)
        w = torch.randn(cout, cin, k, generator=g) / (cin * k) ** 0.5
        ref = F.conv1d(x, w, None, stride=s)                                   # [B, C_out, F]
        Fr = (T - k) // s + 1
        xt = x.transpose(1, 2).contiguous()                                    # time-major [B, T, C_in]
        a = torch.as_strided(xt, (2, Fr, k * cin), (T * cin, s * cin, 1))      # overlapping rows, no copy
        w2 = w.permute(0, 2, 1).reshape(cout, k * cin)                         # [C_out, (tap, channel)]
        out = (a @ w2.t()).transpose(1, 2)
        assert out.shape == ref.shape and (out - ref).abs().max().item() < 1e-4
