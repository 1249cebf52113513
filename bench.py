#!/usr/bin/env python
"""bench.py — SSL train-step samples/sec (ViT-S CIFAR-100, FlexMatch+SemiReward) on N B200s, next to the CPU reference.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--batch-size B] [--stage 1|2]

One "step" = SRFlexMatch.train_step + ParamUpdateHook.after_train_step (backward + AdamW + scheduler) on one batch of
BASELINE.json configs[1] (config/SemiReward/usb_cv/flexmatch/flexmatch_cifar100_200_0.yaml: vit_small_patch2_32,
batch_size 8, uratio 1 -> 24 samples per step and rank, AdamW lr 5e-4 layer_decay 0.5, DropPath 0.2, fp32), stage 1
(0 < it < start_timing: one backbone pass, Rewarder trained on the labelled batch every step) — SURVEY.md §8d.
Synthetic N(0,1) images, random-init weights.  samples/s = world_size * 24 * K / max-over-ranks device time.

Prints ONE JSON line (rank 0).  `value` = inputs already resident in HBM; `e2e` = the same steps driven through the
public API from pinned host tensors (H2D of the batch and D2H of the loss vector inside the timed region).
`--impl reference` times the CPU restatement of the reference's own train_step (oracle/, pinned bit-exactly against the
live reference in the build container; the reference itself is not installable on the GPU box) on the host cores."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

YAML_CFG = dict(  # flexmatch_cifar100_200_0.yaml:10-54 (+ code defaults)
    algorithm="srflexmatch", net="vit_small_patch2_32", optim="AdamW", lr=5e-4, layer_decay=0.5, weight_decay=5e-4,
    num_train_iter=204800, num_warmup_iter=5120, start_timing=20000, N_k=10, batch_size=8, uratio=1, num_classes=100,
    ulb_dest_len=50000, feature_dim=384, sr_lr=5e-4, sr_ema=False, use_cat=True, amp=False, ema_m=0.0, img_size=32,
    p_cutoff=0.95, thresh_warmup=True, ulb_loss_ratio=1.0, clip_grad=0)
F_FWD_GF = 12.134          # GFLOP per sample forward
# BASELINE configs[2] (a parity-test case, not the headline): FreeMatch+SemiReward, vit_base_patch16_224, 1000 classes,
# synthetic 224x224, global batch 1024 over 8 ranks = 128 per GPU (SURVEY.md §8d "Config 3"); `--config 3`, GPU arm only.
YAML_CFG3 = dict(YAML_CFG, algorithm="srfreematch", net="vit_base_patch16_224", num_classes=1000, batch_size=128, img_size=224,
                 feature_dim=768, use_quantile=True, clip_thresh=False, ema_p=0.999, ent_loss_ratio=0.0001, hard_label=True, T=0.5)
F_FWD_GF3 = 35.128
# BASELINE configs[3]: SoftMatch+SemiReward, bert_base_uncased, IMDb-like text: config/usb_nlp/softmatch/softmatch_aclImdb_20_0.yaml
# (2 classes, use_cat False, max_length 512, AdamW lr 5e-5, layer_decay 0.75) + the SR keys of
# config/SemiReward/usb_nlp/softmatch/softmatch_ag_news_40_0.yaml:50-55 (SURVEY.md §8d "Config 4"); `--config 4`.
YAML_CFG4 = dict(algorithm="srsoftmatch", net="bert_base_uncased", optim="AdamW", lr=5e-5, layer_decay=0.75, weight_decay=5e-4,
                 num_train_iter=102400, num_warmup_iter=5120, start_timing=10000, N_k=10, batch_size=8, uratio=1, num_classes=2,
                 ulb_dest_len=25000, feature_dim=768, sr_lr=5e-4, sr_ema=False, use_cat=False, amp=False, ema_m=0.0, max_length=512,
                 ulb_loss_ratio=1.0, clip_grad=0, dist_align=True, dist_uniform=True, ema_p=0.999, n_sigma=2, per_class=False, hard_label=True, T=0.5)
F_FWD_GF4 = 96.64          # GFLOP per sequence forward at L = 512
# BASELINE configs[4]: config/SemiReward/usb_audio/flexmatch/flexmatch_urbansound8k_100_0.yaml unchanged (hubert_base, 10 classes,
# use_cat False, 4 s @ 16 kHz, AdamW lr 5e-5, layer_decay 0.75), 2 ranks (SURVEY.md §8d "Config 5"); `--config 5`.
YAML_CFG5 = dict(algorithm="srflexmatch", net="hubert_base", optim="AdamW", lr=5e-5, layer_decay=0.75, weight_decay=2e-5,
                 num_train_iter=102400, num_warmup_iter=5120, start_timing=10000, N_k=10, batch_size=8, uratio=1, num_classes=10,
                 ulb_dest_len=7000, feature_dim=768, sr_lr=5e-4, sr_ema=False, use_cat=False, amp=False, ema_m=0.0, max_length_seconds=4.0,
                 sample_rate=16000, p_cutoff=0.95, thresh_warmup=True, hard_label=True, T=0.5, ulb_loss_ratio=1.0, clip_grad=0)
# BASELINE configs[0]: config/classic_cv/flexmatch/flexmatch_cifar100_400_0.yaml (wrn_28_2, batch_size 64, uratio 7, SGD lr 0.03 momentum 0.9
# nesterov, weight_decay 1e-3, ema_m 0.999) with `algorithm: srflexmatch` + the SR keys (no SR YAML exists under config/classic_cv; SURVEY.md
# §8d "Config 1"); `--config 1`.  BatchNorm couples the rows: the backward runs over all 960 images (9 B F of algorithmic FLOPs, not 7 B F).
YAML_CFG1 = dict(algorithm="srflexmatch", net="wrn_28_2", optim="SGD", lr=0.03, momentum=0.9, layer_decay=1.0, weight_decay=1e-3,
                 num_train_iter=1048576, num_warmup_iter=0, start_timing=20000, N_k=10, batch_size=64, uratio=7, num_classes=100,
                 ulb_dest_len=50000, feature_dim=128, sr_lr=5e-4, sr_ema=False, use_cat=True, amp=False, ema_m=0.999, img_size=32,
                 p_cutoff=0.95, thresh_warmup=True, hard_label=True, T=0.5, ulb_loss_ratio=1.0, clip_grad=0)
F_FWD_GF1 = 0.429          # GFLOP per 32 x 32 image forward
F_FWD_GF5 = 56.9           # GFLOP per 4 s clip forward (oracle/hubert_oracle.py HubertCfg.fwd_flops_per_clip)
METRICS = {1: "SSL train-step samples/sec (WRN-28-2 CIFAR-100)", 2: "SSL train-step samples/sec (ViT-S CIFAR-100)", 3: "SSL train-step samples/sec (ViT-S CIFAR-100)",
           4: "SSL train-step samples/sec (BERT-base IMDb)", 5: "SSL train-step samples/sec (HuBERT-base UrbanSound8k)"}


def workload_name(config, B, uratio, stage):
    if config == 1:
        return f"srflexmatch wrn_28_2 cifar100 batch_size {B} uratio {uratio} SGD stage {stage} (BASELINE configs[0])"
    if config == 2:
        return f"srflexmatch vit_small_patch2_32 cifar100 batch_size {B} uratio {uratio} stage {stage} (BASELINE configs[1])"
    if config == 3:
        return f"srfreematch vit_base_patch16_224 synthetic 224x224 1000 classes batch_size {B} per GPU stage {stage} (BASELINE configs[2])"
    if config == 5:
        return (f"srflexmatch hubert_base 64000-sample clips (4 s @ 16 kHz) 10 classes batch_size {B} uratio {uratio} stage {stage}, dropout 0.1 at every site, "
                f"SpecAugment on, LayerDrop off (BASELINE configs[4])")
    return f"srsoftmatch bert_base_uncased max_length 512 2 classes batch_size {B} uratio {uratio} stage {stage}, padding tails, dropout 0.1 (BASELINE configs[3])"


def config_block(config, B, uratio, stage, world, setup_steps):
    """The `config` object of the JSON line; both arms (native, reference) print the same keys."""
    per = B * (1 + 2 * uratio)
    f = {1: F_FWD_GF1, 2: F_FWD_GF, 3: F_FWD_GF3, 4: F_FWD_GF4, 5: F_FWD_GF5}[config]
    rows = B * (1 + 2 * uratio)
    gflop = rows * 3 * f if config == 1 else B * (3 + 4 * uratio) * f   # WRN: BatchNorm couples the rows, all of them are back-propagated
    return dict(workload=workload_name(config, B, uratio, stage), samples_per_step_per_gpu=per, parallelism=f"dp{world}",
                drop_path=0.2 if config in (2, 3) else None, dropout=0.1 if config in (4, 5) else None, setup_steps=setup_steps,
                arithmetic="fp32 semantics: bf16x3 split-precision tcgen05 MMA, fp32 accumulate",
                l2=("step working set (~1.8 GB of activations at batch 8) >> 126 MB L2; rotating input batches" if config not in (4, 5) else
                    "step working set (~7-8 GB of activations at batch 8) >> 126 MB L2; rotating input batches"),
                launch="CUDA-graph replay of the backbone forward/backward (SRW_GRAPHS) + programmatic dependent launch (SRW_PDL); "
                       "backward launched inside train_step ahead of the loss read-back",
                algorithmic_gflop_per_step_per_gpu=gflop)   # forward on B (1 + 2u) rows + backward (2x) on the B (1 + u) gradient rows


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch-size", type=int, default=8, help="per-GPU labelled batch (config-faithful: 8)")
    ap.add_argument("--stage", type=int, default=1, choices=[1, 2])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configs index + 1: 1 = WRN-28-2 CIFAR-100 (SGD), 2 = headline ViT-S CIFAR-100, 3 = ViT-B/16 224 FreeMatch, "
                         "4 = BERT-base text, 5 = HuBERT-base audio")
    ap.add_argument("--no-eager-leg", action="store_true", help="skip the informational torch-eager fp32 leg on the same GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._p, self._t = index, [], None, None

    def _run(self):
        # one long-lived `nvidia-smi -lms 100` (spawning it per sample costs ~0.5 s and yields 2 samples per second of bench)
        try:
            self._p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self._p.stdout:
                line = line.strip()
                if line:
                    self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        time.sleep(0.3)   # let the first sample arrive before the timed region starts
        return self

    def __exit__(self, *a):
        if self._p is not None:
            try:
                self._p.terminate()
            except Exception:
                pass
        self._t.join(timeout=3)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons), samples=len(self.rows))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1383.8), d.get("bf16_tflops", 1635.7), d.get("hbm_gbs", 6483.3), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------
# CPU reference arm
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(cfg, steps, warmup, stage):
    """Stage-1 (or stage-2) steps of the oracle = CPU restatement of the reference's train_step + ParamUpdateHook, fp32,
    all host threads; identical synthetic tensors and deterministic weights.  Returns (samples/s, seconds/step, cores)."""
    import torch
    from oracle import ssl_oracle as O
    from semireward_b200 import detgen
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vc = O.ViTConfig(depth=12, num_classes=cfg["num_classes"], drop_path_rate=0.2)
    sc = O.StepConfig(algorithm="srflexmatch", num_classes=cfg["num_classes"], ulb_dest_len=cfg["ulb_dest_len"], p_cutoff=cfg["p_cutoff"],
                      thresh_warmup=cfg["thresh_warmup"], start_timing=cfg["start_timing"], N_k=cfg["N_k"], num_train_iter=cfg["num_train_iter"],
                      num_warmup_iter=cfg["num_warmup_iter"], lr=cfg["lr"], weight_decay=cfg["weight_decay"], layer_decay=cfg["layer_decay"],
                      sr_lr=cfg["sr_lr"], feature_dim=cfg["feature_dim"])
    orc = O.build_det_oracle(vc, sc, seed=0)
    orc.drop_gen = torch.Generator().manual_seed(0)
    it0 = 1 if stage == 1 else cfg["start_timing"] + 1 + 8 * cfg["num_train_iter"]  # K = 8 in stage 2
    times = []
    for i in range(warmup + steps):
        b = O.to_torch_batch(detgen.ssl_batch(cfg["batch_size"], cfg["uratio"], cfg["num_classes"], cfg["ulb_dest_len"], seed=1, step=i))
        t0 = time.perf_counter()
        orc.train_step(b, it0 + i)
        orc.param_update()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    samples = cfg["batch_size"] * (1 + 2 * cfg["uratio"])
    return samples / per_step, per_step, cores


def cpu_reference_run_other(config, steps, warmup):
    """CPU timing of the oracles of the BASELINE configs that have no CUDA path yet: configs[0] WRN-28-2
    (oracle/wrn_oracle.py), configs[3] BERT-base (oracle/bert_oracle.py), configs[4] HuBERT-base (oracle/hubert_oracle.py), all
    pinned against the live reference.  Bounded samples of the named workloads (stated in the returned description)."""
    import torch
    from oracle import ssl_oracle as O
    from semireward_b200 import detgen
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    common = dict(ulb_dest_len=50000, start_timing=20000, N_k=10, num_train_iter=204800, num_warmup_iter=0, sr_lr=5e-4)
    if config == 1:     # config/classic_cv/flexmatch/flexmatch_cifar100_400_0.yaml + SR keys: B 64, uratio 7, SGD
        from oracle import wrn_oracle as WO
        B, u, C = 64, 7, 100
        sc = O.StepConfig(algorithm="srflexmatch", num_classes=C, lr=0.03, weight_decay=1e-3, layer_decay=1.0, feature_dim=128, **common)
        orc = WO.build_det_wrn_oracle(WO.WRNCfg(num_classes=C), sc, seed=0)
        batch = lambda i: O.to_torch_batch(detgen.ssl_batch(B, u, C, 50000, seed=1, step=i))
        name, what = "WRN-28-2 CIFAR-100", f"srflexmatch wrn_28_2 cifar100 batch_size {B} uratio {u} SGD stage 1 (BASELINE configs[0])"
    elif config == 4:   # config/usb_nlp/softmatch/softmatch_aclImdb_20_0.yaml + SR keys: max_length 512, 2 classes; sample: batch 2 of 8
        from oracle import bert_oracle as BO
        B, u, C = 2, 1, 2
        sc = O.StepConfig(algorithm="srsoftmatch", num_classes=C, lr=5e-5, weight_decay=5e-4, layer_decay=0.75, feature_dim=768, **common)
        orc = BO.build_det_bert_oracle(BO.BertCfg(num_classes=C, hidden_dropout=0.0, attn_dropout=0.0, pooled_dropout=0.0), sc, seed=0)

        def batch(i):
            b = detgen.nlp_batch(B, u, C, 50000, max_length=512, seed=1, step=i)
            return {k: ({kk: torch.from_numpy(vv) for kk, vv in v.items()} if isinstance(v, dict) else torch.from_numpy(v)) for k, v in b.items()}
        name, what = "BERT-base IMDb", f"srsoftmatch bert_base_uncased max_length 512 batch_size {B} (of 8) uratio {u} stage 1, dropout off (BASELINE configs[3])"
    else:               # config/SemiReward/usb_audio/flexmatch/flexmatch_urbansound8k_100_0.yaml: 4 s clips, 10 classes; sample: batch 2 of 8
        from oracle import hubert_oracle as HO
        B, u, C = 2, 1, 10
        sc = O.StepConfig(algorithm="srflexmatch", num_classes=C, lr=2e-5, weight_decay=5e-4, layer_decay=0.75, feature_dim=768, **common)
        orc = HO.build_det_hubert_oracle(HO.HubertCfg(num_classes=C), sc, seed=0)
        batch = lambda i: O.to_torch_batch(detgen.audio_batch(B, u, C, 50000, samples=64000, seed=1, step=i))
        name, what = "HuBERT-base UrbanSound8k", f"srflexmatch hubert_base 64000-sample clips batch_size {B} (of 8) uratio {u} stage 1, dropout / LayerDrop / SpecAugment off (BASELINE configs[4])"
    times = []
    for i in range(warmup + steps):
        b = batch(i)
        t0 = time.perf_counter()
        orc.train_step(b, 1 + i)
        orc.param_update()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    return B * (1 + 2 * u) / per_step, per_step, cores, name, what, B * (1 + 2 * u)


def cpu_reference_bert(B, u, steps, warmup, dropout):
    """BASELINE configs[3] on the host cores: the oracle restatement of ClassificationBert + SRSoftMatch step (pinned against the
    live reference and Hugging Face BertModel, tests/test_bert_oracle.py), bert-base, L = 512.  -> (samples/s, s/step, cores)."""
    import torch
    from oracle import bert_oracle as BO, ssl_oracle as O
    from semireward_b200 import detgen
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c = YAML_CFG4
    sc = O.StepConfig(algorithm="srsoftmatch", num_classes=c["num_classes"], lr=c["lr"], weight_decay=c["weight_decay"], layer_decay=c["layer_decay"],
                      feature_dim=768, ulb_dest_len=c["ulb_dest_len"], start_timing=c["start_timing"], N_k=c["N_k"], num_train_iter=c["num_train_iter"],
                      num_warmup_iter=c["num_warmup_iter"], sr_lr=c["sr_lr"])
    p = dropout
    orc = BO.build_det_bert_oracle(BO.BertCfg(num_classes=c["num_classes"], hidden_dropout=p, attn_dropout=p, pooled_dropout=p), sc, seed=0, stochastic=p > 0)
    orc.drop_gen = torch.Generator().manual_seed(0)
    times = []
    for i in range(warmup + steps):
        b = detgen.nlp_batch(B, u, c["num_classes"], c["ulb_dest_len"], max_length=512, seed=1, step=i)
        b = {k: ({kk: torch.from_numpy(vv) for kk, vv in v.items()} if isinstance(v, dict) else torch.from_numpy(v)) for k, v in b.items()}
        t0 = time.perf_counter()
        orc.train_step(b, 1 + i)
        orc.param_update()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    return B * (1 + 2 * u) / per_step, per_step, cores


def cpu_reference_hubert(B, u, steps, warmup, dropout):
    """BASELINE configs[4] on the host cores: the oracle restatement of ClassificationHubert + SRFlexMatch step (pinned against the live
    reference and Hugging Face HubertModel, tests/test_hubert_oracle.py), hubert-base, 4 s clips.  -> (samples/s, s/step, cores)."""
    import torch
    from oracle import hubert_oracle as HO, ssl_oracle as O
    from semireward_b200 import detgen
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c = YAML_CFG5
    sc = O.StepConfig(algorithm="srflexmatch", num_classes=c["num_classes"], lr=c["lr"], weight_decay=c["weight_decay"], layer_decay=c["layer_decay"],
                      feature_dim=768, ulb_dest_len=c["ulb_dest_len"], start_timing=c["start_timing"], N_k=c["N_k"], num_train_iter=c["num_train_iter"],
                      num_warmup_iter=c["num_warmup_iter"], sr_lr=c["sr_lr"], p_cutoff=c["p_cutoff"], thresh_warmup=c["thresh_warmup"])
    p = dropout
    orc = HO.build_det_hubert_oracle(HO.HubertCfg(num_classes=c["num_classes"], feat_proj_dropout=p, hidden_dropout=p, attention_dropout=p,
                                                  activation_dropout=p, pooled_dropout=p), sc, seed=0)
    orc.drop_seed = 0 if p > 0 else None
    times = []
    for i in range(warmup + steps):
        b = O.to_torch_batch(detgen.audio_batch(B, u, c["num_classes"], c["ulb_dest_len"], samples=64000, seed=1, step=i))
        t0 = time.perf_counter()
        orc.train_step(b, 1 + i)
        orc.param_update()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    return B * (1 + 2 * u) / per_step, per_step, cores


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    note = "one CPU process on the host cores whatever --gpus says (the reference's CPU path does not shard over GPUs)"
    if a.config == 1:
        steps, warmup = max(1, min(a.steps, 2)), max(0, min(a.warmup, 1))     # tens of seconds per CPU step
        sps, per_step, cores, name, what, samples = cpu_reference_run_other(a.config, steps, warmup)
        B1 = YAML_CFG1["batch_size"]
        print(json.dumps(dict(impl="reference", metric=METRICS[1], value=sps, unit="samples/s", n_gpus=a.gpus, steps=steps,
                              warmup=warmup, ms_per_step=per_step * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                              data="synthetic", config=config_block(1, B1, YAML_CFG1["uratio"], a.stage, a.gpus, SETUP_STEPS), note=note,
                              cpu_baseline=dict(value=sps, unit="samples/s", cores=cores, kind="port",
                                                sample=f"{steps} step(s) of the oracle restatement (pinned against the live reference), {warmup} warm-up"),
                              e2e=dict(value=sps, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return
    if a.config == 4:
        steps, warmup = max(1, min(a.steps, 2)), max(0, min(a.warmup, 1))
        B = a.batch_size
        sps, per_step, cores = cpu_reference_bert(B, 1, steps, warmup, 0.1)
        sample = f"{steps} stage-1 step(s) of the oracle restatement at the full batch (dropout 0.1 from torch's CPU generator), {warmup} warm-up"
    elif a.config == 5:
        steps, warmup = max(1, min(a.steps, 2)), max(0, min(a.warmup, 1))
        sps, per_step, cores = cpu_reference_hubert(a.batch_size, 1, steps, warmup, 0.1)
        sample = (f"{steps} stage-1 step(s) of the oracle restatement at the full batch (counter-based dropout 0.1 at every site; LayerDrop and SpecAugment off), "
                  f"{warmup} warm-up")
    else:
        cfg = dict(YAML_CFG, batch_size=a.batch_size)
        steps, warmup = max(1, min(a.steps, 8)), max(0, min(a.warmup, 2))   # bounded sample: ~1 s per CPU step, at most 8 + 2 steps
        if a.config == 3:
            raise SystemExit("bench.py --impl reference --config 3: the CPU arm of ViT-B/16 224 at batch 128 is not bounded to minutes; use --config 2")
        sps, per_step, cores = cpu_reference_run(cfg, steps, warmup, a.stage)
        sample = f"{steps} stage-{a.stage} steps of the oracle restatement (bit-exact vs the live reference), {warmup} warm-up"
    line = dict(impl="reference", metric=METRICS[a.config], value=sps, unit="samples/s", n_gpus=a.gpus,
                steps=steps, warmup=warmup, ms_per_step=per_step * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", config=config_block(a.config, a.batch_size, 1, a.stage, a.gpus, SETUP_STEPS), note=note,
                cpu_baseline=dict(value=sps, unit="samples/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=sps, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# informational leg: the same step in PyTorch eager fp32 on the SAME GPU (TF32 off) — what the reference's own modules would run
# as on a B200 (cuBLAS sgemm, unfused attention).  Context for the roofline numbers, never the headline or a baseline.
# ------------------------------------------------------------------------------------------------
def torch_eager_fp32_leg(cfg, stage, steps=6, warmup=2):
    import torch
    from oracle import ssl_oracle as O
    from semireward_b200 import detgen
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda")
    vc = O.ViTConfig(depth=12, num_classes=cfg["num_classes"], drop_path_rate=0.2)
    sc = O.StepConfig(algorithm="srflexmatch", num_classes=cfg["num_classes"], ulb_dest_len=cfg["ulb_dest_len"], p_cutoff=cfg["p_cutoff"],
                      thresh_warmup=cfg["thresh_warmup"], start_timing=cfg["start_timing"], N_k=cfg["N_k"], num_train_iter=cfg["num_train_iter"],
                      num_warmup_iter=cfg["num_warmup_iter"], lr=cfg["lr"], weight_decay=cfg["weight_decay"], layer_decay=cfg["layer_decay"],
                      sr_lr=cfg["sr_lr"], feature_dim=cfg["feature_dim"])
    orc = O.build_det_oracle(vc, sc, seed=0)
    for d in (orc.p, orc.rp, orc.gp):
        for k in d:
            d[k] = d[k].detach().to(dev).requires_grad_(True)
    orc.opt, orc.ropt, orc.gopt = O.AdamState(orc.p, decoupled=True), O.AdamState(orc.rp, decoupled=False), O.AdamState(orc.gp, decoupled=False)
    orc.hook.selected_label, orc.hook.classwise_acc = orc.hook.selected_label.to(dev), orc.hook.classwise_acc.to(dev)
    orig = O.draw_drop_path_masks
    O.draw_drop_path_masks = lambda c, b, g=None: (lambda m: None if m is None else m.to(dev))(orig(c, b, g))
    try:
        it0 = 1 if stage == 1 else cfg["start_timing"] + 1 + 8 * cfg["num_train_iter"]
        batches = [{k: v.to(dev) for k, v in O.to_torch_batch(detgen.ssl_batch(cfg["batch_size"], cfg["uratio"], cfg["num_classes"], cfg["ulb_dest_len"], seed=1, step=i)).items()}
                   for i in range(4)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(warmup + steps):
            if i == warmup:
                torch.cuda.synchronize()
                e0.record()
            orc.train_step(dict(batches[i % 4]), it0 + i)
            orc.param_update()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    finally:
        O.draw_drop_path_masks = orig
    samples = cfg["batch_size"] * (1 + 2 * cfg["uratio"])
    return dict(value=samples / (ms * 1e-3), unit="samples/s", ms_per_step=ms, steps=steps,
                what="functional torch restatement of the reference step (oracle/ssl_oracle.py) run on cuda:0 in eager fp32, TF32 off: cuBLAS sgemm, "
                     "unfused attention, torch autograd, per-tensor AdamW, host-side hook bookkeeping — informational only")


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
SETUP_STEPS = 12


def main_native(a):
    import torch
    import torch.distributed as dist
    import semireward_b200 as S
    from semireward_b200 import _lib as L, detgen
    from semireward_b200.parallel import send_model_cuda

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's "NCCL version ..." banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = L.load()
    L.check(lib.srw_device_check(None, None, None), "srw_device_check")

    base_cfg = {1: YAML_CFG1, 2: YAML_CFG, 3: YAML_CFG3, 4: YAML_CFG4, 5: YAML_CFG5}[a.config]
    bs = a.batch_size if (a.batch_size != 8 or a.config not in (1, 3)) else base_cfg["batch_size"]
    cfg = dict(base_cfg, batch_size=bs, gpu=local, distributed=world > 1, world_size=world, rank=rank)
    args = S.get_config(cfg)
    torch.manual_seed(0)
    builder = S.get_net_builder(args.net, False)
    if a.config == 5:   # LayerDrop off: the timed step runs all 12 layers for every clip (never less work than the reference's expectation)
        import functools
        builder = functools.partial(builder, layerdrop=0.0)
    alg = S.get_algorithm(args, builder, None, None)
    alg.model = send_model_cuda(args, alg.model)
    alg.model.train()
    it0 = 1 if a.stage == 1 else args.start_timing + 1 + 8 * args.num_train_iter
    B, U = args.batch_size, args.batch_size * args.uratio
    samples_per_step = B + 2 * U
    text = a.config == 4
    audio = a.config == 5

    def host_batch(i):
        if audio:
            b = detgen.audio_batch(B, args.uratio, args.num_classes, args.ulb_dest_len, samples=int(args.max_length_seconds * args.sample_rate), seed=1 + rank, step=i)
            return {k: torch.from_numpy(v).pin_memory() for k, v in b.items()}
        if text:
            b = detgen.nlp_batch(B, args.uratio, args.num_classes, args.ulb_dest_len, max_length=args.max_length, seed=1 + rank, step=i)
            return {k: ({kk: torch.from_numpy(vv).pin_memory() for kk, vv in v.items()} if isinstance(v, dict) else torch.from_numpy(v).pin_memory())
                    for k, v in b.items()}
        b = detgen.ssl_batch(B, args.uratio, args.num_classes, args.ulb_dest_len, img_size=args.img_size, seed=1 + rank, step=i)
        return {k: torch.from_numpy(v).pin_memory() for k, v in b.items()}

    def to_dev(v):
        return {k: t.cuda(non_blocking=True) for k, t in v.items()} if isinstance(v, dict) else v.cuda(non_blocking=True)

    def nbytes(v):
        return sum(nbytes(t) for t in v.values()) if isinstance(v, dict) else v.numel() * v.element_size()

    n_batches = 4 if a.config != 3 else 2   # config 3: 231 MB per host batch
    hbatches = [host_batch(i) for i in range(n_batches)]
    import inspect
    step_keys = set(inspect.signature(alg.train_step).parameters)   # process_batch's filter (algorithmbase.py:287-296)
    dbatches = [{k: to_dev(v) for k, v in hb.items() if k in step_keys} for hb in hbatches]
    h2d = sum(nbytes(v) for k, v in hbatches[0].items() if k in step_keys)

    def step_device(i):
        alg.it = it0 + i
        alg.out_dict, alg.log_dict = alg.train_step(**dbatches[i % n_batches])
        alg.call_hook("after_train_step", "ParamUpdateHook")   # the metric excludes the EMA / logging hooks

    def step_e2e(i):
        alg.it = it0 + i
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**hbatches[i % n_batches]))
        alg.call_hook("after_train_step", "ParamUpdateHook")   # the metric excludes the EMA / logging hooks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, offset):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            fn(offset + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    W, K = max(3, a.warmup), a.steps
    if (text or audio or a.config == 1) and a.steps == 300:
        K = 60          # a BERT / HuBERT step is ~10x a ViT-S step: keep the default run within minutes
    # one-time engine set-up outside the measurement (reported as config.setup_steps): the first call of a backbone pass runs
    # eagerly, the second is captured into a CUDA graph, kernels are lazily loaded on first use and stage 2 alternates between
    # step variants (SR update every N_k steps) — with a short --warmup those one-offs would land in the timed region
    for i in range(SETUP_STEPS):
        step_device(i)
    for i in range(W):
        step_device(i)
    launches0 = lib.srw_kernel_launches()
    with ClockSampler(local) as clk:
        ms = timed(step_device, K, W)
    launches = lib.srw_kernel_launches() - launches0
    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(step_e2e, K, W + K)
    value = world * samples_per_step * K / (ms / 1e3)
    e2e = world * samples_per_step * K / (ms_e2e / 1e3)
    step_ms = ms / K

    roof = None
    if not a.no_roofline:   # every rank runs the profiled steps (they contain the gradient all-reduce); rank 0 reports
        lib.srw_profile_enable(1)
        st = (L.ProfileStats * L.PROF_NUM)()
        nprof = min(K, 5)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0.record()
        for i in range(nprof):
            step_device(W + 2 * K + i)
        t1.record(); torch.cuda.synchronize()
        step_ms_prof = t0.elapsed_time(t1) / nprof
        L.check(lib.srw_profile_collect(st), "srw_profile_collect")
        lib.srw_profile_enable(0)
        sust, burst, hbm, how = peaks()
        g = st[L.PROF_GEMM]
        ach = g.flops / (g.total_ms * 1e-3) / 1e12 if g.total_ms > 0 else 0.0
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
        if os.path.isfile(tp) and a.config == 2:   # dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch from the committed `ncu --set full` capture
            tj = json.load(open(tp))
            traffic, traffic_src = tj.get("gemm_mean_traffic_bytes"), tj.get("source")

        def kern(stat, peak, unit="TFLOP/s"):
            if stat.total_ms <= 0:
                return None
            v = (stat.flops if unit == "TFLOP/s" else stat.bytes) / (stat.total_ms * 1e-3) / (1e12 if unit == "TFLOP/s" else 1e9)
            # share_of_step: the kernel class's device time per step over the TIMED (graph-replayed) step; the per-launch events are
            # taken in a separate pass with graph replay off (profiled_step_ms), kernel durations are the same kernels' either way
            return dict(bound="tensor" if unit == "TFLOP/s" else "hbm", achieved=v, peak=peak, unit=unit, frac=v / peak,
                        avg_launch_us=1e3 * stat.total_ms / max(stat.launches, 1), launches_per_step=stat.launches / nprof,
                        share_of_step=stat.total_ms / nprof / step_ms)
        roof = kern(g, sust)
        roof.update(kernel="gemm_bf16x3_tcgen05_kernel", traffic=traffic, traffic_source=traffic_src,
                    algorithmic_bytes_per_launch=g.bytes / max(g.launches, 1), peak_source=f"MEASURED_PEAKS.json bf16_tflops_sustained ({how})",
                    profiled_step_ms=step_ms_prof,
                    note="achieved = algorithmic 2MNK per launch / CUDA-event launch time; the kernel issues 3 bf16 MMAs per algorithmic "
                         "product (hi*hi, hi*lo, lo*hi) for fp32-level accuracy, so the tensor pipe does 3x these FLOPs (frac is capped at 1/3)")
        # the attention kernels (north_star quotes the attention-GEMM roofline) and the optimizer, as siblings of the dominant kernel
        roof["attention_fwd"] = kern(st[L.PROF_ATTN_FWD], sust)
        roof["attention_bwd"] = kern(st[L.PROF_ATTN_BWD], sust)
        roof["adamw"] = kern(st[L.PROF_ADAMW], hbm, "GB/s")

    cpu = eager = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline and a.config == 2:
        sps, per_step, cores = cpu_reference_run(dict(YAML_CFG, batch_size=a.batch_size), 6, 1, a.stage)
        cpu = dict(value=sps, unit="samples/s", cores=cores, kind="port",
                   sample=f"6 stage-{a.stage} steps (+1 warm-up) of the oracle restatement of the reference train_step+ParamUpdateHook, "
                          f"{per_step:.2f} s/step, torch CPU fp32, {cores} threads")
    if rank == 0 and world == 1 and not a.no_cpu_baseline and a.config == 4:
        sps, per_step, cores = cpu_reference_bert(2, 1, 1, 0, 0.1)
        cpu = dict(value=sps, unit="samples/s", cores=cores, kind="port",
                   sample=f"1 stage-1 step of the oracle restatement at batch_size 2 (of {B}; 6 sequences of 512 tokens), {per_step:.1f} s/step, torch CPU fp32, {cores} threads")
    if rank == 0 and world == 1 and not a.no_cpu_baseline and a.config == 1:
        sps, per_step, cores, _, what, _ = cpu_reference_run_other(1, 1, 0)
        cpu = dict(value=sps, unit="samples/s", cores=cores, kind="port",
                   sample=f"1 stage-1 step of the oracle restatement at the full batch (960 images), {per_step:.1f} s/step, torch CPU fp32, {cores} threads")
    if rank == 0 and world == 1 and not a.no_cpu_baseline and a.config == 5:
        sps, per_step, cores = cpu_reference_hubert(2, 1, 1, 0, 0.1)
        cpu = dict(value=sps, unit="samples/s", cores=cores, kind="port",
                   sample=f"1 stage-1 step of the oracle restatement at batch_size 2 (of {B}; 6 clips of 64000 samples), {per_step:.1f} s/step, torch CPU fp32, {cores} threads")
    if rank == 0 and world == 1 and not a.no_eager_leg and a.config == 2:
        try:
            eager = torch_eager_fp32_leg(dict(YAML_CFG, batch_size=a.batch_size), a.stage)
        except Exception as e:   # informational leg: never fail the bench on it
            eager = dict(unavailable=f"{type(e).__name__}: {e}"[:200])
    if rank == 0:
        line = dict(metric=METRICS[a.config], value=value, unit="samples/s", n_gpus=world, steps=K, warmup=W,
                    ms_per_step=step_ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=config_block(a.config, B, args.uratio, a.stage, world, SETUP_STEPS),
                    clocks=clk.summary(), gpu_launches=int(launches),
                    e2e=dict(value=e2e, unit="samples/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4 * (5 + U), ms_per_step=ms_e2e / K))
        if roof is not None:
            line["roofline"] = roof
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if eager is not None:
            line["torch_eager_fp32_b200"] = eager
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    import faulthandler
    # a hung collective or kernel must end as a traceback + non-zero exit, not as a silent stall of the whole run
    faulthandler.dump_traceback_later(int(os.environ.get("SRW_BENCH_WATCHDOG", "900")), exit=True)
    args_ = parse()
    if args_.impl == "reference":
        main_reference(args_)
    else:
        main_native(args_)
