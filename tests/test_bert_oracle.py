"""CPU: the text-path oracle (oracle/bert_oracle.py: ClassificationBert + the SSL step with use_cat False, SURVEY.md §8a
row a4 / BASELINE configs[3]) against golden vectors generated from the LIVE reference by tests/golden/make_golden_bert.py
and — when /root/reference is mounted (build container) — against the live reference and Hugging Face BertModel themselves.
The reference ships no fixtures for this path; these files are the pin.

Bars: forward bit-exact against transformers' eager attention, <= 2e-6 against its default (sdpa) kernels; step losses equal
to float32 round-off; parameters after AdamW within 1e-4 (Adam turns noise-level gradient entries — e.g. the key bias,
whose gradient is mathematically zero — into +-lr moves, so two fp32 implementations separate by O(lr) in those entries)."""
import os

import numpy as np
import pytest
import torch

from golden_cases import BERT_CASES, BERT_SMALL, STEPS, bert_small_cfg

GOLD = os.path.join(os.path.dirname(__file__), "golden")
LIVE = os.path.isdir("/root/reference/semilearn")


def _bert_cfg(cfg, **over):
    from oracle import bert_oracle as BO
    kw = dict(vocab_size=BERT_SMALL["vocab_size"], layers=BERT_SMALL["layers"], max_position=BERT_SMALL["max_position"], hidden_dropout=0.0,
              attn_dropout=0.0, pooled_dropout=0.0, num_classes=cfg["num_classes"])
    kw.update(over)
    return BO.BertCfg(**kw)


def _step_cfg(cfg):
    from oracle import ssl_oracle as O
    return O.StepConfig(algorithm=cfg["algorithm"], num_classes=cfg["num_classes"], ulb_dest_len=cfg["ulb_dest_len"], p_cutoff=cfg["p_cutoff"],
                        thresh_warmup=cfg["thresh_warmup"], start_timing=cfg["start_timing"], N_k=cfg["N_k"], num_train_iter=cfg["num_train_iter"],
                        num_warmup_iter=cfg["num_warmup_iter"], lr=cfg["lr"], weight_decay=cfg["weight_decay"], layer_decay=cfg["layer_decay"],
                        sr_lr=cfg["sr_lr"], feature_dim=cfg["feature_dim"])


def _text_batch(cfg, it):
    from semireward_b200 import detgen
    b = detgen.nlp_batch(cfg["batch_size"], cfg["uratio"], cfg["num_classes"], cfg["ulb_dest_len"], max_length=BERT_SMALL["max_length"],
                         vocab_size=BERT_SMALL["vocab_size"], seed=1, step=it)
    return {k: ({kk: torch.from_numpy(vv) for kk, vv in v.items()} if isinstance(v, dict) else torch.from_numpy(v)) for k, v in b.items()}


@pytest.mark.parametrize("name", sorted(BERT_CASES))
def test_bert_oracle_matches_golden(name):
    from oracle import bert_oracle as BO
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    spec = BERT_CASES[name]
    cfg = bert_small_cfg(**spec["cfg"])
    orc = BO.build_det_bert_oracle(_bert_cfg(cfg), _step_cfg(cfg), seed=0, head_gain=spec["head_gain"])
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    utils = []
    for it in range(STEPS):
        rec = orc.train_step(_text_batch(cfg, it), it)
        feat_lb = rec["feat_lb"].numpy().copy()
        orc.param_update()
        for key, val in (("loss", rec["total_loss"]), ("sup_loss", rec["sup_loss"]), ("unsup_loss", rec["unsup_loss"])):
            np.testing.assert_allclose(np.float32(float(val)), gold[f"it{it}_{key}"], rtol=2e-5, atol=2e-6, err_msg=f"it{it} {key}")
        assert abs(float(rec["util_ratio"]) - float(gold[f"it{it}_util_ratio"])) < 1e-6, it
        utils.append(float(rec["util_ratio"]))
        np.testing.assert_allclose(feat_lb, gold[f"it{it}_feat_lb"], rtol=0, atol=2e-4 if it else 1e-6, err_msg=f"it{it} feat_lb")
        for key, pname, row in (("cls_bias", "classifier.2.bias", None), ("q0_row0", "bert.encoder.layer.0.attention.self.query.weight", 0),
                                ("word_row7", "bert.embeddings.word_embeddings.weight", 7)):
            v = orc.p[pname].detach().numpy()
            np.testing.assert_allclose(v if row is None else v[row], gold[f"it{it}_{key}"], rtol=0, atol=1e-4, err_msg=f"it{it} {key}")
        rsum = sum(v.detach().double().sum().item() for v in orc.rp.values())
        assert abs(rsum - float(gold[f"it{it}_rewarder_sum"])) < 5e-2, (it, rsum, float(gold[f"it{it}_rewarder_sum"]))
        if cfg["algorithm"] == "srsoftmatch":
            assert abs(float(orc.hook.prob_max_mu_t) - float(gold[f"it{it}_mu"])) < 1e-6
            assert abs(float(orc.hook.prob_max_var_t) - float(gold[f"it{it}_var"])) < 1e-6
    if name == "bert_srfixmatch_l2":
        assert any(0.0 < u < 1.0 for u in utils), utils   # the fixture really exercises a mixed mask
    # the first update is the cleanest check of the optimizer table: everything but noise-level entries moved identically
    assert np.abs(orc.p["classifier.2.bias"].detach().numpy() - gold[f"it{STEPS - 1}_cls_bias"]).max() < 1e-4


def test_bert_known_answers():
    """Frozen from reference probes: FLOP count, optimizer table, padding semantics."""
    from oracle import bert_oracle as BO
    assert abs(BO.BertCfg().fwd_flops_per_seq(512) / 1e9 - 96.64) < 0.05
    c = BO.BertCfg()
    shapes = c.param_shapes()
    assert len(shapes) == 5 + 16 * 12 + 2 + 4 and sum(int(np.prod(s)) for _, s in shapes) == 109482240 + 768 * 768 + 768 + 2 * 768 + 2
    hp = BO.bert_param_hparams(shapes, 12, 5e-5, 5e-4, 0.75)
    assert hp["classifier.2.weight"] == (5e-5, 5e-4) and hp["bert.pooler.dense.bias"] == (5e-5, 0.0)
    assert abs(hp["bert.embeddings.word_embeddings.weight"][0] - 5e-5 * 0.75 ** 13) < 1e-18
    assert abs(hp["bert.encoder.layer.11.output.dense.weight"][0] - 5e-5 * 0.75) < 1e-18
    assert len({(round(lr / 5e-5, 12), wd) for lr, wd in hp.values()}) == 28      # 14 layer ids x decay / no-decay
    # padding: a padded key never receives attention weight, and the padding token's embedding row never receives a gradient
    cfg = bert_small_cfg()
    bc = _bert_cfg(cfg)
    from semireward_b200 import detgen
    p = {n: torch.from_numpy(detgen.fill_param(n, s, 0)).requires_grad_(True) for n, s in bc.param_shapes()}
    x = _text_batch(cfg, 0)["x_lb"]
    assert int((x["attention_mask"] == 0).sum()) > 0
    logits, feat = BO.bert_forward(p, x, bc)
    x2 = {k: v.clone() for k, v in x.items()}
    pad = x2["attention_mask"] == 0
    x2["input_ids"][pad] = 5          # tokens under the mask are never attended to, but their OWN positions are still part of the
    logits2, _ = BO.bert_forward(p, x2, bc)   # mean pool (bert.py:36-37 averages all L positions), so the logits do change
    assert not torch.equal(logits, logits2)
    logits.sum().backward()
    assert float(p["bert.embeddings.word_embeddings.weight"].grad[0].abs().max()) == 0.0
    assert p["bert.pooler.dense.weight"].grad is None


@pytest.mark.skipif(not LIVE, reason="live reference / its transformers pin only exist in the build container")
def test_bert_forward_matches_huggingface_eager_and_sdpa():
    from transformers import BertConfig, BertModel
    from oracle import bert_oracle as BO
    cfg = bert_small_cfg()
    bc = _bert_cfg(cfg)
    x = _text_batch(cfg, 0)["x_lb"]
    for impl, tol in (("eager", 0.0), ("sdpa", 2e-6)):
        hc = BertConfig(vocab_size=bc.vocab_size, num_hidden_layers=bc.layers, max_position_embeddings=bc.max_position, hidden_dropout_prob=0.0,
                        attention_probs_dropout_prob=0.0, attn_implementation=impl)
        torch.manual_seed(0)
        m = BertModel(hc).train()
        p = {"bert." + n: v.detach().clone() for n, v in m.named_parameters()}
        assert list(p) == [n for n, _ in bc.param_shapes() if n.startswith("bert.")]
        p.update({"classifier.0.weight": torch.zeros(768, 768), "classifier.0.bias": torch.zeros(768),
                  "classifier.2.weight": torch.zeros(bc.num_classes, 768), "classifier.2.bias": torch.zeros(bc.num_classes)})
        ref = m(**x, return_dict=True)["last_hidden_state"].mean(1)
        _, feat = BO.bert_forward(p, x, bc)
        err = (feat - ref).abs().max().item()
        assert err <= tol, (impl, err)


@pytest.mark.skipif(not LIVE, reason="live reference only exists in the build container")
def test_bert_oracle_matches_live_reference_default_attention():
    """Same cases against the reference running transformers' DEFAULT attention kernels (sdpa), i.e. exactly what
    `BertModel.from_pretrained(name)` (bert.py:13) would run: losses to 1e-5, optimizer table identical."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden_bert as G
    from oracle import bert_oracle as BO, ref_driver as R
    name = "bert_srfixmatch_l2"
    spec = BERT_CASES[name]
    ref = G.run_case(name, spec, attn="sdpa")
    cfg = bert_small_cfg(**spec["cfg"])
    orc = BO.build_det_bert_oracle(_bert_cfg(cfg), _step_cfg(cfg), seed=0, head_gain=spec["head_gain"])
    for it in range(STEPS):
        rec = orc.train_step(_text_batch(cfg, it), it)
        orc.param_update()
        assert abs(float(rec["total_loss"]) - float(ref[f"it{it}_loss"])) < 1e-5 * max(1.0, abs(float(ref[f"it{it}_loss"]))), it
        assert abs(float(rec["util_ratio"]) - float(ref[f"it{it}_util_ratio"])) < 1e-6, it
    # optimizer table against the live AdamW param groups
    alg = R.build_reference_algorithm(dict(cfg, dist_align=True, dist_uniform=True, n_sigma=2, per_class=False),
                                      net_kwargs=dict(bert=dict(vocab_size=64, num_hidden_layers=2, max_position_embeddings=64), dropout=0.0))
    names = {id(p): n for n, p in alg.model.named_parameters()}
    hp = BO.bert_param_hparams(BO.BertCfg(vocab_size=64, layers=2, max_position=64, num_classes=cfg["num_classes"]).param_shapes(), 2, cfg["lr"],
                               cfg["weight_decay"], cfg["layer_decay"])
    seen = 0
    for g in alg.optimizer.param_groups:
        for p in g["params"]:
            lr, wd = hp[names[id(p)]]
            assert abs(g["lr"] - lr) < 1e-15 and g["weight_decay"] == wd, names[id(p)]
            seen += 1
    assert seen == len(hp)


def test_counter_dropout_masks_are_reproducible_and_unbiased():
    """The counter-based masks (what a native epilogue will regenerate): keep rate, independence across sites, and a stochastic
    forward that is a pure function of (inputs, seed)."""
    from oracle import bert_oracle as BO
    m3 = BO.counter_keep_mask(1 << 18, 0.9, 7, 3)
    assert abs(float(m3.float().mean()) - 0.9) < 3e-3
    assert 0.15 < float((BO.counter_keep_mask(1 << 18, 0.9, 7, 4) != m3).float().mean()) < 0.21     # 2 * 0.9 * 0.1 for independent sites
    assert torch.equal(BO.counter_keep_mask(1000, 0.9, 7, 3), m3[:1000])                            # element i does not depend on the extent
    # scalar restatement of the same 32-bit hash (include/srw.h: srw_dropout; csrc/srw_common.cuh: lowbias32 / drop_kept)
    def mix(x):
        x &= 0xFFFFFFFF
        x ^= x >> 16
        x = (x * 0x7FEB352D) & 0xFFFFFFFF
        x ^= x >> 15
        x = (x * 0x846CA68B) & 0xFFFFFFFF
        return x ^ (x >> 16)
    keep, key, site = 0.9, 7, 3
    sk = mix(key + site * 0x9E3779B9)
    want = [(mix(i + sk) >> 8) < int(keep * (1 << 24)) for i in range(512)]
    assert want == m3[:512].tolist()
    assert torch.equal(BO.counter_keep_mask(100, 0.9, 7, 3, offset=400), m3[400:500])
    assert BO.call_key(5, 0) != BO.call_key(5, 1) != BO.call_key(6, 1)
    cfg = bert_small_cfg()
    bc = _bert_cfg(cfg, hidden_dropout=0.1, attn_dropout=0.1, pooled_dropout=0.1)
    from semireward_b200 import detgen
    p = {n_: torch.from_numpy(detgen.fill_param(n_, s, 0)) for n_, s in bc.param_shapes()}
    x = _text_batch(cfg, 0)["x_lb"]
    a, _ = BO.bert_forward(p, x, bc, BO.BertDropout(bc, 11, True))
    b, _ = BO.bert_forward(p, x, bc, BO.BertDropout(bc, 11, True))
    c, _ = BO.bert_forward(p, x, bc, BO.BertDropout(bc, 12, True))
    d, _ = BO.bert_forward(p, x, bc)
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, d)


def test_two_sweep_chunked_attention_plan_is_exact():
    """The plan for L = 512: sweep 1 walks the 256-key chunks only for the row max and the row sum (a scalar
    rescale when the max moves), sweep 2 recomputes the scores, forms P = exp(S - max) as split bf16 and accumulates P V with
    no rescale of O.  Emulated here with the split-plane products (hi*hi + hi*lo + lo*hi) and an additive key-padding mask,
    against plain fp32 softmax attention."""
    def split(x):
        hi = x.to(torch.bfloat16).float()
        return hi, (x - hi).to(torch.bfloat16).float()

    def mm3(a, b):
        ah, al = split(a)
        bh, bl = split(b)
        return ah @ bh + (ah @ bl + al @ bh)
    g = torch.Generator().manual_seed(0)
    L, dh, chunk = 512, 64, 256
    q, k, v = (torch.randn(3, L, dh, generator=g) * s for s in (1.5, 1.5, 1.0))
    lens = torch.tensor([512, 300, 129])
    keymask = torch.zeros(3, 1, L).masked_fill(torch.arange(L)[None, None, :] >= lens[:, None, None], torch.finfo(torch.float32).min)
    ref = torch.softmax(q @ k.transpose(1, 2) * dh ** -0.5 + keymask, -1) @ v
    scale = dh ** -0.5
    m = torch.full((3, L), -float("inf"))
    ssum = torch.zeros(3, L)
    for c in range(0, L, chunk):                                   # sweep 1: statistics only
        s = mm3(q, k[:, c:c + chunk].transpose(1, 2)) * scale + keymask[:, :, c:c + chunk]
        m_new = torch.maximum(m, s.max(-1)[0])
        ssum = ssum * torch.exp(m - m_new) + torch.exp(s - m_new[..., None]).sum(-1)
        m = m_new
    o = torch.zeros(3, L, dh)
    for c in range(0, L, chunk):                                   # sweep 2: P in place, PV accumulate, no rescale
        s = mm3(q, k[:, c:c + chunk].transpose(1, 2)) * scale + keymask[:, :, c:c + chunk]
        o = o + mm3(torch.exp(s - m[..., None]), v[:, c:c + chunk])
    out = o / ssum[..., None]
    # scores reach +-20 with these operands, so the 2^-17 relative error of a split-plane product is ~1e-4 absolute inside exp()
    assert (out - ref).abs().max().item() < 1e-4
    lse = m + torch.log(ssum)
    assert (lse - torch.logsumexp(q @ k.transpose(1, 2) * scale + keymask, -1)).abs().max().item() < 1e-4
