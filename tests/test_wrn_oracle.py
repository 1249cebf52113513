"""CPU: the WRN-path oracle (oracle/wrn_oracle.py: WideResNet + use_cat True + SGD, SURVEY.md §8a row a5 / BASELINE
configs[0]) against golden vectors generated from the LIVE reference by tests/golden/make_golden_wrn.py and — when
/root/reference is mounted (build container) — bit for bit against the live reference itself, BatchNorm buffers included.
The reference ships no fixtures for this path; these files are the pin.  Same torch CPU primitives on both
sides (F.conv2d, F.batch_norm), so the bar is float32 round-off; integers and masks bit-exact."""
import os

import numpy as np
import pytest
import torch

from golden_cases import STEPS, WRN_CASES, wrn_small_cfg

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
    b = detgen.ssl_batch(cfg["batch_size"], cfg["uratio"], cfg["num_classes"], cfg["ulb_dest_len"], seed=1, step=it)
    return {k: torch.from_numpy(v) for k, v in b.items()}


@pytest.mark.parametrize("name", sorted(WRN_CASES))
def test_wrn_oracle_matches_golden(name):
    from oracle import wrn_oracle as WO
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    spec = WRN_CASES[name]
    cfg = wrn_small_cfg(**spec["cfg"])
    orc = WO.build_det_wrn_oracle(WO.WRNCfg(depth=spec["depth"], num_classes=cfg["num_classes"]), _step_cfg(cfg), seed=0, head_gain=spec["head_gain"])
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    bn_w0 = orc.p["block3.layer.0.bn1.weight"].detach().numpy().copy()
    utils = []
    for it in range(STEPS):
        rec = orc.train_step(_batch(cfg, it), it)
        orc.param_update()
        got = {"loss": rec["total_loss"], "sup_loss": rec["sup_loss"], "unsup_loss": rec["unsup_loss"], "util_ratio": rec["util_ratio"]}
        for k, v in got.items():
            np.testing.assert_allclose(np.float32(float(v)), gold[f"it{it}_{k}"], rtol=2e-5, atol=2e-6, err_msg=f"it{it} {k}")
        utils.append(float(rec["util_ratio"]))
        for key, val in (("cls_bias", orc.p["classifier.bias"]), ("conv1_w0", orc.p["conv1.weight"][0]), ("b3_bn1_weight", orc.p["block3.layer.0.bn1.weight"]),
                         ("b3_bn1_mean", orc.buf["block3.layer.0.bn1.running_mean"]), ("bn1_var", orc.buf["bn1.running_var"])):
            np.testing.assert_allclose(val.detach().numpy(), gold[f"it{it}_{key}"], rtol=2e-5, atol=2e-6, err_msg=f"it{it} {key}")
        psum = sum(v.detach().double().sum().item() for v in orc.p.values()) + sum(v.double().sum().item() for v in orc.buf.values())
        np.testing.assert_allclose(psum, float(gold[f"it{it}_param_sum"]), rtol=1e-6, err_msg=f"it{it} param_sum")
        rsum = sum(v.detach().double().sum().item() for v in orc.rp.values())
        np.testing.assert_allclose(rsum, float(gold[f"it{it}_rewarder_sum"]), rtol=1e-6, err_msg=f"it{it} rewarder_sum")
        if cfg["algorithm"] == "srflexmatch":
            assert np.array_equal(orc.hook.selected_label.numpy(), gold[f"it{it}_selected_label"]), it
            assert np.array_equal(orc.hook.classwise_acc.numpy(), gold[f"it{it}_classwise_acc"]), it
    # wrn.py:46-54 quirk: block3.layer.0.bn1 is evaluated (its running mean moves away from 0) but its output is unused, so its
    # scale never receives a gradient and SGD never touches it
    assert np.array_equal(orc.p["block3.layer.0.bn1.weight"].detach().numpy(), bn_w0)
    assert float(orc.buf["block3.layer.0.bn1.running_mean"].abs().max()) > 0.0
    if name == "wrn_srfixmatch_d10":
        assert any(0.0 < u < 1.0 for u in utils), utils


def test_wrn_known_answers():
    from oracle import wrn_oracle as WO
    c = WO.WRNCfg()
    shapes = c.param_shapes()
    assert len(shapes) == 81 and sum(int(np.prod(s)) for _, s in shapes) == 1479236                 # live probe of wrn_28_2(num_classes=100)
    assert abs(c.fwd_flops_per_image() / 1e9 - 0.429) < 1e-3
    assert len(c.bn_names()) == 25
    hp = WO.wrn_param_hparams(shapes, 0.03, 1e-3)
    assert hp["conv1.weight"] == (0.03, 1e-3) and hp["conv1.bias"] == (0.03, 0.0) and hp["block2.layer.1.bn2.weight"] == (0.03, 0.0)
    assert hp["classifier.weight"] == (0.03, 1e-3) and sum(1 for _, wd in hp.values() if wd > 0) == 1 + 24 + 3 + 1
    # BatchNorm couples the rows: changing one weak row changes the logits of the labelled rows
    p = {n: torch.randn(s, generator=torch.Generator().manual_seed(i)) * 0.1 + (1.0 if ("bn" in n and n.endswith("weight")) else 0.0) for i, (n, s) in enumerate(WO.WRNCfg(depth=10).param_shapes())}
    cfg10 = WO.WRNCfg(depth=10)
    x = torch.randn(6, 3, 32, 32, generator=torch.Generator().manual_seed(0))
    l1, _ = WO.wrn_forward(p, WO.new_bn_buffers(cfg10), x, cfg10)
    x2 = x.clone()
    x2[5] += 1.0
    l2, _ = WO.wrn_forward(p, WO.new_bn_buffers(cfg10), x2, cfg10)
    assert not torch.equal(l1[0], l2[0])
    # SGD with nesterov momentum, first two steps by hand: buf1 = g, p1 = p - lr (g + mu g); buf2 = mu g + g2 ...
    sgd = WO.SGDState({"w": None})
    w = {"w": torch.tensor([1.0])}
    sgd.step(w, {"w": torch.tensor([0.5])}, {"w": (0.1, 0.0)})
    assert abs(w["w"].item() - (1.0 - 0.1 * (0.5 + 0.9 * 0.5))) < 1e-7
    sgd.step(w, {"w": torch.tensor([0.5])}, {"w": (0.1, 0.0)})
    assert abs(w["w"].item() - (0.905 - 0.1 * (0.5 + 0.9 * (0.9 * 0.5 + 0.5)))) < 1e-7


@pytest.mark.skipif(not LIVE, reason="live reference only exists in the build container")
def test_wrn_oracle_bit_exact_against_live_reference_depth28():
    """The full WRN-28-2 of configs[0] for three steps (stage 1 with an SR update): parameters, BatchNorm buffers and Rewarder
    bit for bit, optimizer table against the live SGD param groups."""
    import inspect
    from oracle import ref_driver as R, wrn_oracle as WO
    cfg = wrn_small_cfg(batch_size=2, uratio=1)
    alg = R.build_reference_algorithm(dict(cfg))
    R.load_det_weights(alg, seed=0, head_gain=4.0)
    wc = WO.WRNCfg(num_classes=cfg["num_classes"])
    assert [(n, tuple(p.shape)) for n, p in alg.model.named_parameters()] == wc.param_shapes()
    orc = WO.build_det_wrn_oracle(wc, _step_cfg(cfg), seed=0, head_gain=4.0)
    names = {id(p): n for n, p in alg.model.named_parameters()}
    for g in alg.optimizer.param_groups:
        assert g["momentum"] == 0.9 and g["nesterov"] is True
        for p in g["params"]:
            assert (g["lr"], g["weight_decay"]) == orc.hp[names[id(p)]], names[id(p)]
    for it in range(3):
        b = _batch(cfg, it)
        alg.it = it
        rb = {k: v for k, v in b.items() if k in inspect.signature(alg.train_step).parameters}
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**rb))
        alg.hooks_dict["ParamUpdateHook"].after_train_step(alg)
        rec = orc.train_step(b, it)
        orc.param_update()
        assert abs(alg.log_dict["train/total_loss"] - float(rec["total_loss"])) < 1e-6
        for n, p in alg.model.named_parameters():
            assert torch.equal(p.detach(), orc.p[n].detach()), (it, n)
        for n, v in alg.model.named_buffers():
            if not n.endswith("num_batches_tracked"):
                assert torch.equal(v, orc.buf[n]), (it, n)
        for n, p in alg.rewarder.named_parameters():
            assert torch.equal(p.detach(), orc.rp[n].detach()), (it, n)
