"""CPU: the C-ABI library loads and exports every symbol include/srw.h declares (no compute calls), and the host-side
mirror of the reference's plugin surface behaves like the reference's (names, argument meaning, errors)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as G
    G.build()
    from semireward_b200 import _lib as L
    lib = L.load()
    hdr = open(os.path.join(ROOT, "include", "srw.h")).read()
    declared = set(re.findall(r"\b(srw_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"srw_epilogue", "srw_gemm_impl", "srw_profile_class"}
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/srw.h but not exported"
    bound = {n for n, _, _ in L.SYMBOLS}
    assert declared <= bound, f"not bound in _lib.py: {sorted(declared - bound)}"
    assert lib.srw_version() >= 1
    # same number of parameters in every prototype and in its ctypes argtypes
    hdr_nc = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    arity = {m.group(1): (0 if m.group(2).strip() in ("", "void") else m.group(2).count(",") + 1)
             for m in re.finditer(r"\b(srw_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", hdr_nc)}
    assert len(arity) >= 30
    for name, _, args in L.SYMBOLS:
        assert arity[name] == len(args), f"{name}: {arity[name]} parameters in include/srw.h, {len(args)} in _lib.py"


def test_struct_sizes_match_header_layout():
    """ctypes mirrors of the argument structs: spot-check sizes that would drift if a field were added on one side."""
    import ctypes as C
    from semireward_b200 import _lib as L
    assert C.sizeof(L.AdamWRow) == 5 * 8 + 8 + 8 + 4 + 4 + 8 + 8 + 8
    assert C.sizeof(L.ProfileStats) == 32
    assert C.sizeof(L.VitConfig) == 9 * 4


def test_every_struct_matches_the_header_field_by_field(tmp_path):
    """include/srw.h compiled as plain C11 by gcc: sizeof and every offsetof against the ctypes mirrors in _lib.py (a field
    added, renamed or retyped on one side only fails here, on CPU, instead of corrupting arguments on the GPU box)."""
    import ctypes as C
    import subprocess
    from semireward_b200 import _lib as L
    pairs = dict(srw_profile_stats="ProfileStats", srw_split_args="SplitArgs", srw_gemm_args="GemmArgs", srw_splitk_reduce_args="SplitKReduceArgs",
                 srw_colsum_args="ColsumArgs", srw_fold_colsum="FoldColsum", srw_grad_fold_args="GradFoldArgs", srw_layernorm_fwd_args="LayerNormFwdArgs",
                 srw_layernorm_bwd_args="LayerNormBwdArgs", srw_attn_fwd_args="AttnFwdArgs", srw_attn_bwd_args="AttnBwdArgs", srw_vit_config="VitConfig",
                 srw_vit_fwd_args="VitFwdArgs", srw_vit_bwd_args="VitBwdArgs", srw_rewarder_fwd_args="RewarderFwdArgs",
                 srw_generator_fwd_args="GeneratorFwdArgs", srw_rewarder_train_args="RewarderTrainArgs", srw_flexmatch_mask_args="FlexMatchMaskArgs",
                 srw_ssl_loss_args="SslLossArgs", srw_freematch_mask_args="FreeMatchMaskArgs", srw_freematch_entropy_args="FreeMatchEntropyArgs",
                 srw_softmatch_mask_args="SoftMatchMaskArgs", srw_adamw_row="AdamWRow", srw_adamw_args="AdamWArgs", srw_ema_row="EmaRow",
                 srw_ema_args="EmaArgs", srw_dropout="Dropout", srw_bert_config="BertConfig", srw_bert_fwd_args="BertFwdArgs",
                 srw_bert_bwd_args="BertBwdArgs", srw_hubert_config="HubertConfig", srw_hubert_fwd_args="HubertFwdArgs",
                     srw_hubert_bwd_args="HubertBwdArgs", srw_wrn_config="WrnConfig", srw_wrn_fwd_args="WrnFwdArgs", srw_wrn_bwd_args="WrnBwdArgs",
                     srw_sgd_args="SgdArgs", srw_aug_op_desc="AugOpDesc", srw_aug_sample="AugSample", srw_augment_args="AugmentArgs")
    hdr = open(os.path.join(ROOT, "include", "srw.h")).read()
    assert set(re.findall(r"}\s*(srw_[a-z0-9_]+);", hdr)) == set(pairs), "a struct of include/srw.h has no ctypes mirror in this table"
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "srw.h"', "int main(void) {"]
    for cname, pyname in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for f in getattr(L, pyname)._fields_:
            lines.append(f'  printf("{cname} {f[0]} %zu\\n", offsetof({cname}, {f[0]}));')
    lines += ["  return 0;", "}"]
    src, exe = tmp_path / "abi.c", tmp_path / "abi"
    src.write_text("\n".join(lines))
    r = subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:2000]
    out = subprocess.run([str(exe)], capture_output=True, text=True).stdout.strip().split("\n")
    assert len(out) > 300
    for ln in out:
        cname, field, val = ln.split()
        cls = getattr(L, pairs[cname])
        exp = C.sizeof(cls) if field == "size" else getattr(cls, field).offset
        assert int(val) == exp, f"{cname}.{field}: header {val}, ctypes {exp}"


def test_registry_and_config_surface():
    import semireward_b200 as S
    assert "srflexmatch" in S.ALGORITHMS and "srflexmatch" in S.name2alg.keys()
    args = S.get_config(dict(algorithm="srflexmatch", ulb_dest_len=100))
    assert args.p_cutoff == 0.95 and args.start_timing == 20000 and args.feature_dim == 384 and args.N_k == 10 and args.thresh_warmup is True
    with pytest.raises(KeyError, match="Unknown algorithm"):
        S.get_config(dict(algorithm="nope"))
    args.algorithm = "nope"
    with pytest.raises(KeyError, match="Unknown algorithm"):
        S.get_algorithm(args, None, None, None)
    with pytest.raises(KeyError):
        S.get_net_builder("resnet_nope")
    assert S.get_net_builder("vit_small_patch2_32").__name__ == "vit_small_patch2_32"


def test_net_builder_state_dict_contract():
    import semireward_b200 as S
    from oracle import ssl_oracle as O
    m = S.get_net_builder("vit_small_patch2_32")(num_classes=100)
    sd = m.state_dict()
    ref = O.ViTConfig().param_shapes()
    assert [k for k in sd] == [n for n, _ in ref] and len(sd) == 152
    assert all(tuple(sd[n].shape) == tuple(s) for n, s in ref)
    assert m.num_features == 384 and m.no_weight_decay() == {"pos_embed", "cls_token"}
    feat = torch.randn(5, 384)
    assert torch.equal(m(feat, only_fc=True), torch.nn.functional.linear(feat, m.head.weight, m.head.bias))   # vit.py:293-294
    assert set(m.group_matcher()) == {"stem", "blocks"}
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 32, 32))   # no CPU fallback


def test_param_groups_match_reference_table():
    import semireward_b200 as S
    from oracle import ssl_oracle as O
    from semireward_b200.core.optim import get_cosine_schedule_with_warmup, get_optimizer
    m = S.get_net_builder("vit_small_patch2_32")(num_classes=100)
    opt = get_optimizer(m, "AdamW", 5e-4, 0.9, 5e-4, 0.5)
    assert len(opt.param_groups) == 28
    hp = O.vit_param_hparams(O.ViTConfig().param_shapes(), 12, 5e-4, 5e-4, 0.5)
    names = {id(p): n for n, p in m.named_parameters()}
    for g in opt.param_groups:
        for p in g["params"]:
            lr, wd = hp[names[id(p)]]
            assert abs(g["lr"] - lr) < 1e-18 and g["weight_decay"] == wd, names[id(p)]
    sched = get_cosine_schedule_with_warmup(opt, 204800, num_warmup_steps=5120)
    assert sched.get_last_lr()[0] == 0.0
    assert abs(O.cosine_lr_factor(5120, 204800, 5120) - 1.0) < 1e-12 and O.cosine_lr_factor(100, 204800, 5120) == 100 / 5120


def test_algorithm_construction_and_process_batch_filter():
    import functools
    import semireward_b200 as S
    args = S.get_config(dict(algorithm="srflexmatch", optim="AdamW", lr=5e-4, layer_decay=0.5, ulb_dest_len=64, sr_ema=False, num_train_iter=64))
    if torch.cuda.is_available():
        pytest.skip("CPU-only construction test")
    alg = S.get_algorithm(args, functools.partial(S.get_net_builder(args.net), depth=1), None, None)
    assert list(alg.hooks_dict) == ["ParamUpdateHook", "EMAHook", "PseudoLabelingHook", "MaskingHook"]   # priority order (algorithmbase.py:226-240)
    assert alg.registered_hook("MaskingHook") and not alg.registered_hook("DistAlignHook")
    import inspect
    assert list(inspect.signature(alg.train_step).parameters) == ["x_lb", "y_lb", "idx_ulb", "x_ulb_w", "x_ulb_s"]
    alg.it = 20001
    assert alg.sr_decay() == int(max(8, 1 + 64 / 20001))
    sd = alg.get_save_dict()
    assert {"model", "ema_model", "optimizer", "scheduler", "it", "epoch", "best_it", "best_eval_acc", "classwise_acc", "selected_label"} <= set(sd)
    assert sd["selected_label"].shape == (64,) and int(sd["selected_label"][0]) == -1
    # the SemiReward state travels under one extra key (the reference's checkpoints drop it); a round trip restores it,
    # and a reference-style checkpoint without the key still loads
    sr = sd["semireward"]
    assert set(sr) == {"rewarder", "generator", "rewarder_adam"} and sr["rewarder_adam"] is None and len(sr["rewarder"]) == 17
    saved = {k: v.clone() for k, v in sr["rewarder"].items()}
    with torch.no_grad():
        for p_ in alg.rewarder.parameters():
            p_.add_(1.0)
    ps = alg.rewarder._params()
    adam = dict(m=[torch.full_like(t, 0.5) for t in ps], v=[torch.full_like(t, 0.25) for t in ps], step=7)
    alg._sr_load(dict(semireward=dict(rewarder=saved, generator=sr["generator"], rewarder_adam=adam)))
    assert all(torch.equal(v, saved[k]) for k, v in alg.rewarder.state_dict().items())
    st = alg.rewarder.optimizer_state()
    assert st["step"] == 7 and float(st["m"][0].flatten()[0]) == 0.5 and float(st["v"][-1].flatten()[0]) == 0.25
    alg._sr_load({})
    for name in ("srfreematch", "srsoftmatch", "srfixmatch", "srpseudolabel"):
        a2 = S.get_algorithm(S.get_config(dict(algorithm=name, optim="AdamW", lr=5e-4, layer_decay=0.5, ulb_dest_len=64, sr_ema=False, num_train_iter=64)),
                             functools.partial(S.get_net_builder(args.net), depth=1), None, None)
        assert "semireward" in a2.get_save_dict(), name


def test_detgen_is_platform_independent():
    """Counter-based integer hashing + exact float64 arithmetic: fixed checksums on every machine."""
    import hashlib
    from semireward_b200 import detgen
    x = detgen.normal("x_lb", (4, 3, 8, 8), 7)
    assert hashlib.sha256(x.tobytes()).hexdigest()[:16] == hashlib.sha256(detgen.normal("x_lb", (4, 3, 8, 8), 7).tobytes()).hexdigest()[:16]
    assert x.dtype.name == "float32" and abs(float(x.mean())) < 0.2 and 0.7 < float(x.std()) < 1.3
    idx = detgen.distinct_integers("idx_ulb", 8, 50000, 3)
    assert len(set(idx.tolist())) == 8 and idx.min() >= 0 and idx.max() < 50000
    golden = os.path.join(ROOT, "tests", "golden", "detgen_checksums.txt")
    lines = []
    for tag, shape, seed in (("x_lb", (8, 3, 32, 32), 1000003), ("head.weight", (100, 384), 0), ("rewarder.label_embedding.weight", (100, 128), 0)):
        arr = detgen.normal(tag, shape, seed) if tag == "x_lb" else detgen.fill_param(tag, shape, seed)
        lines.append(f"{tag} {hashlib.sha256(arr.tobytes()).hexdigest()}")
    if not os.path.isfile(golden):
        open(golden, "w").write("\n".join(lines) + "\n")
    assert open(golden).read().split("\n")[:3] == lines


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the arm the driver runs next to ours) on the host cores: one JSON line with the contract keys.
    Also the cheapest guard against a syntax / import error in bench.py before it reaches the GPU box."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().split("\n") if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == "SSL train-step samples/sec (ViT-S CIFAR-100)" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0
    # the native arm never falls back to anything: without a GPU it fails loudly (every config has a CUDA path and only a CUDA path) ...
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "5", "--steps", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert r.returncode != 0 and not [ln for ln in r.stdout.split("\n") if ln.startswith("{")], r.stdout[-500:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/semilearn"), reason="live reference only exists in the build container")
def test_integration_recipe_combined_class_constructs_on_the_reference_base():
    """`class SRFlexMatch(Native, RefBase)` — native step and hooks, the reference's loop / dataset methods."""
    if torch.cuda.is_available():
        pytest.skip("CPU-only construction test")
    from oracle import ref_driver as R
    R.load_reference()
    from semilearn.core import AlgorithmBase as RefBase
    from semilearn.lighting.config import get_config
    import semireward_b200 as S
    from semireward_b200.algorithms.srflexmatch import SRFlexMatch as Native

    class Combined(Native, RefBase):
        pass
    cfg = dict(R.DEFAULT_CFG)
    ulb = cfg.pop("ulb_dest_len")
    cfg.pop("drop_path")
    args = get_config(cfg)
    args.ulb_dest_len = ulb
    alg = Combined(args, S.get_net_builder("vit_small_patch2_32"), None, None)
    assert type(alg.model).__module__ == "semireward_b200.nets.vit"
    assert Combined.train_step is Native.train_step and Combined.set_hooks is Native.set_hooks
    for name in ("train", "set_dataset", "set_data_loader", "evaluate"):
        assert getattr(Combined, name) is getattr(RefBase, name), name
    assert list(alg.hooks_dict)[:2] == ["ParamUpdateHook", "EMAHook"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/semilearn"), reason="live reference only exists in the build container")
def test_combined_class_steps_through_the_reference_train_loop(tmp_path, monkeypatch):
    """INTEGRATION.md §1 end to end on the host side: `reference_algorithm(Native)` runs under the reference's own
    `AlgorithmBase.train()` (algorithmbase.py:346-375) with its Evaluation / Checkpoint / Timer / Logging hooks for two
    iterations, `evaluate()` works, the checkpoint it writes is loaded by the native class AND by the unmodified reference
    class.  No GPU here, so only the device work is stubbed (train_step's arithmetic, optimizer / EMA launches, the model
    forward inside evaluate, CUDA events); every call of the loop protocol is the real one."""
    if torch.cuda.is_available():
        pytest.skip("host-protocol test for the CPU container")
    from oracle import ref_driver as R
    R.load_reference()
    from semilearn.core import AlgorithmBase as RefBase
    from semilearn.lighting.config import get_config
    import semireward_b200 as S
    from semireward_b200.core import hooks as H
    from semireward_b200.integration import reference_algorithm, register_into_reference
    from semireward_b200.algorithms.srflexmatch import SRFlexMatch as Native

    class FakeEvent:
        def __init__(self, enable_timing=False): pass
        def record(self): pass
        def elapsed_time(self, other): return 0.0
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(H.EMA, "update", lambda self: None)
    Combined = reference_algorithm(Native)
    assert Combined.train is RefBase.train and Combined.evaluate is RefBase.evaluate and Combined.train_step is Native.train_step
    cfg = dict(R.DEFAULT_CFG, num_train_iter=2, epoch=1, num_eval_iter=2, num_log_iter=1, save_dir=str(tmp_path), save_name="run",
               resume=False, num_warmup_iter=0, start_timing=100)
    ulb = cfg.pop("ulb_dest_len")
    cfg.pop("drop_path")
    args = get_config(cfg)
    args.ulb_dest_len = ulb
    import functools
    builder = functools.partial(S.get_net_builder("vit_small_patch2_32"), depth=1)
    alg = Combined(args, builder, None, None)
    names = list(alg.hooks_dict)
    for h in ("ParamUpdateHook", "EMAHook", "EvaluationHook", "CheckpointHook", "DistSamplerSeedHook", "TimerHook", "LoggingHook",
              "PseudoLabelingHook", "MaskingHook"):
        assert h in names, (h, names)
    assert names.index("ParamUpdateHook") < names.index("EMAHook") < names.index("EvaluationHook") < names.index("CheckpointHook") < names.index("LoggingHook")
    calls = []

    def train_step(x_lb, y_lb, idx_ulb, x_ulb_w, x_ulb_s):   # same parameter names: process_batch filters the batch by them
        calls.append(x_lb.shape[0])
        loss = sum((p * 0).sum() for p in alg.model.parameters()) + 1.0
        return dict(loss=loss, feat={}), {"train/sup_loss": 1.0, "train/unsup_loss": 0.0, "train/total_loss": 1.0, "train/util_ratio": 1.0}
    alg.train_step = train_step
    alg.optimizer.step = lambda *a, **k: None
    alg.model.forward = lambda x, **k: {"logits": torch.zeros(x.shape[0], 100), "feat": torch.zeros(x.shape[0], 384)}
    g = torch.Generator().manual_seed(0)
    lb = [dict(idx_lb=torch.arange(8), x_lb=torch.randn(8, 3, 32, 32, generator=g), y_lb=torch.randint(0, 100, (8,), generator=g)) for _ in range(2)]
    ulb_l = [dict(idx_ulb=torch.arange(8), x_ulb_w=torch.randn(8, 3, 32, 32, generator=g), x_ulb_s=torch.randn(8, 3, 32, 32, generator=g)) for _ in range(2)]
    alg.loader_dict = {"train_lb": lb, "train_ulb": ulb_l, "eval": [dict(x_lb=lb[0]["x_lb"], y_lb=lb[0]["y_lb"])]}
    alg.train()                                          # the reference's loop
    assert calls == [8, 8] and alg.it == 2
    assert "eval/top-1-acc" in alg.log_dict and "train/prefetch_time" in alg.log_dict and "lr" in alg.log_dict
    assert alg.results_dict["eval/best_it"] == alg.best_it
    ck_path = os.path.join(str(tmp_path), "run", "latest_model.pth")
    ck = torch.load(ck_path, map_location="cpu")
    for k in ("model", "ema_model", "optimizer", "loss_scaler", "scheduler", "it", "epoch", "best_it", "best_eval_acc", "classwise_acc",
              "selected_label", "semireward"):
        assert k in ck, k
    assert ck["loss_scaler"]          # non-empty: the reference's enabled GradScaler refuses an empty state (algorithmbase.py:505)
    # native resume: counters, scheduler position and (here empty) optimizer state come back
    alg2 = Combined(args, builder, None, None)
    alg2.load_model(ck_path)
    assert alg2.it == ck["it"] and alg2.scheduler.last_epoch == alg.scheduler.last_epoch == 2
    # the unmodified reference loads the native checkpoint (same keys; 28 AdamW groups; FlexMatch hook state)
    ref = R.build_reference_algorithm(dict(num_train_iter=2, epoch=1, num_warmup_iter=0), net_kwargs=dict(depth=1))
    ref.load_model(ck_path)
    assert ref.it == ck["it"] and ref.scheduler.last_epoch == 2
    for (n1, p1), (n2, p2) in zip(ref.model.state_dict().items(), alg.model.state_dict().items()):
        assert n1 == n2 and torch.equal(p1, p2)
    # and the registry route of INTEGRATION.md: the reference's get_algorithm now builds the combined native class
    import semilearn
    import semilearn.nets as ref_nets
    from semilearn.core.utils import ALGORITHMS as REF
    saved_algs, saved_nets = dict(REF._dict), dict(vars(ref_nets))
    try:
        reg = register_into_reference()
        a3 = semilearn.get_algorithm(args, semilearn.get_net_builder("vit_small_patch2_32", False), None, None)
        assert isinstance(a3, Native) and isinstance(a3, RefBase) and type(a3) is reg["srflexmatch"]
        assert type(a3.model).__module__ == "semireward_b200.nets.vit"
    finally:   # the swap is process-wide: later live-reference tests must get the reference's own classes back
        REF._dict.clear()
        REF._dict.update(saved_algs)
        for k, v in saved_nets.items():
            if getattr(ref_nets, k, None) is not v:
                setattr(ref_nets, k, v)


def test_pretrained_loading_follows_the_reference_load_checkpoint(tmp_path):
    """nets/utils.py:18-73 semantics: weights under 'model', `module.` prefix stripped, classifier tensors skipped,
    pos_embed resampled bicubically to the model's grid, non-strict.  Against the live reference function when it is mounted."""
    import functools
    import semireward_b200 as S
    from semireward_b200.nets.utils import load_checkpoint, resize_pos_embed_vit
    torch.manual_seed(0)
    src = S.get_net_builder("vit_small_patch2_32")(num_classes=10, depth=1, img_size=16)       # 8 x 8 grid + cls
    sd = {("module." + k): v.clone() for k, v in src.state_dict().items()}
    sd["module.fc.weight"] = torch.zeros(3, 3)                                                   # foreign classifier tensors are dropped
    path = os.path.join(str(tmp_path), "pre.pth")
    torch.save({"model": sd}, path)
    dst = S.get_net_builder("vit_small_patch2_32")(num_classes=100, pretrained=True, pretrained_path=path, depth=1)   # 16 x 16 grid
    assert dst.pos_embed.shape == (1, 257, 384)
    want = resize_pos_embed_vit(src.pos_embed.data, dst.pos_embed.data)
    assert torch.equal(dst.pos_embed.data, want)
    assert torch.equal(dst.pos_embed.data[:, 0], src.pos_embed.data[:, 0])                       # the cls slot is copied, the grid resampled
    assert torch.equal(dst.blocks[0].attn.qkv.weight, src.blocks[0].attn.qkv.weight)
    assert dst.head.weight.shape == (100, 384) and not torch.equal(dst.head.bias, torch.full((100,), 7.0))
    if os.path.isdir("/root/reference/semilearn"):
        from oracle import ref_driver as R
        R.load_reference()
        from semilearn.nets.utils import load_checkpoint as ref_load, resize_pos_embed_vit as ref_resize
        assert torch.equal(ref_resize(src.pos_embed.data, dst.pos_embed.data), want)
        ref_dst = S.get_net_builder("vit_small_patch2_32")(num_classes=100, depth=1)
        ref_load(ref_dst, path)
        for (n, a), (_, b) in zip(ref_dst.state_dict().items(), dst.state_dict().items()):
            if not n.startswith("head"):
                assert torch.equal(a, b), n


def test_wrn_and_hubert_builders_param_groups_and_host_randomness():
    """Host logic of the two round-2 backbones without a GPU: state_dict contracts, optimizer groups against the oracles' tables, the
    per-call LayerDrop / SpecAugment bookkeeping of the HuBERT wrapper (what the engine later receives as explicit inputs)."""
    import numpy as np
    import semireward_b200 as S
    from oracle import hubert_oracle as HO, wrn_oracle as WO
    from semireward_b200.core.optim import FusedAdamW, FusedSGD, get_optimizer
    # WRN-28-2: 81 parameters + 75 BatchNorm buffers, SGD groups = (no decay: bn + biases | decay), the two dead bn1 excluded from the gradient list
    wrn = S.get_net_builder("wrn_28_2")(num_classes=100)
    wc = WO.WRNCfg(num_classes=100)
    assert [(n, tuple(p.shape)) for n, p in wrn.named_parameters()] == wc.param_shapes() and len(wrn.state_dict()) == 156
    opt = get_optimizer(wrn, "SGD", 0.03, 0.9, 1e-3, 1.0)
    assert isinstance(opt, FusedSGD) and opt.param_groups[0]["nesterov"] is True
    hp = WO.wrn_param_hparams(wc.param_shapes(), 0.03, 1e-3)
    names = {id(p): n for n, p in wrn.named_parameters()}
    for g in opt.param_groups:
        for p in g["params"]:
            assert (g["lr"], g["weight_decay"]) == hp[names[id(p)]], names[id(p)]
    assert len(wrn._grad_params()) == 77 and wrn.num_features == 128 and not wrn.stochastic()
    with pytest.raises(RuntimeError, match="CUDA"):
        wrn.eval()(torch.zeros(1, 3, 32, 32))
    # HuBERT: HF key names, AdamW layer-decay groups, frames, per-call draws
    hub = S.get_net_builder("hubert_base")(num_classes=10, num_hidden_layers=2)
    hc = HO.HubertCfg(layers=2, num_classes=10)
    assert list(hub.state_dict()) == [n for n, _ in hc.param_shapes()]
    opt = get_optimizer(hub, "AdamW", 5e-5, 0.9, 2e-5, 0.75)
    assert isinstance(opt, FusedAdamW)
    hp = HO.hubert_param_hparams(hc.param_shapes(), 2, 5e-5, 2e-5, 0.75)
    names = {id(p): n for n, p in hub.named_parameters()}
    for g in opt.param_groups:
        for p in g["params"]:
            lr, wd = hp[names[id(p)]]
            assert abs(g["lr"] - lr) < 1e-18 and g["weight_decay"] == wd, names[id(p)]
    assert hub.frames(64000) == hc.frames(64000) == 199 and hub.frames(8000) == 24
    hub.train()
    assert hub.stochastic()
    first = hub.draw_streams(2, 3, 4, "cpu")                      # two passes = six model calls
    assert first == 0 and hub._calls == 6 and sorted(hub._call_draws) == list(range(6))
    spec = hub.streams_for(first, [(0, "lb"), (1, "s"), (0, "w")], 3, 4, "cpu")
    assert spec["segments"].tolist() == [0, 3, 7, 11] and spec["skip"].shape == (3, 2) and spec["keys"].shape == (11,)
    assert spec["rows"].tolist() == [0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3]
    assert len(set(spec["keys"].tolist())) == 3                   # one dropout stream per model call
    assert sorted(hub._call_draws) == [1, 3, 5]                   # the draws of the three calls of this launch were consumed
    m = hub._mask_time(spec, 199)
    assert m.shape == (11, 199) and m.dtype == torch.uint8
    per_clip = m.sum(1)
    assert int(per_clip.min()) >= 10 and int(per_clip.max()) <= 20   # >= 2 spans of 10 frames each (mask_time_min_masks), overlaps merge
    assert torch.equal(m, hub._mask_time(spec, 199))              # drawn from the call's own seed: reproducible for the backward
    hub.set_call_draws(7, layer_skip=[1, 0], mask_time=np.zeros((4, 199), dtype=bool))
    assert hub._call_draws[7][0].tolist() == [1, 0]
    hub.eval()
    assert not hub.stochastic()
