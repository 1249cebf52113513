"""Generate golden vectors from the LIVE reference (build container only; /root/reference must be mounted).

    python tests/golden/make_golden.py

Drives the reference's own SRFlexMatch / SRFreeMatch / SRSoftMatch train_step + ParamUpdateHook on CPU through
oracle/ref_driver.py (import shims only, nothing copied), with the deterministic weights and batches of
semireward_b200.detgen, and records per-step outputs into tests/golden/<case>.npz.  The oracle (oracle/ssl_oracle.py) is
then pinned against these files by tests/test_oracle_golden.py on any machine, and the CUDA path is compared with the
oracle by the -m gpu tests."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden_cases import CASES, STEPS  # noqa: E402
from helpers import batch_tensors, small_cfg  # noqa: E402


def run_case(name, spec):
    from oracle import ref_driver as R
    cfg = small_cfg(**spec["cfg"])
    ref_cfg = {k: v for k, v in cfg.items() if k not in ("gpu",)}
    ref_cfg.update(ema_p=0.999, ent_loss_ratio=0.001, use_quantile=True, clip_thresh=False)
    alg = R.build_reference_algorithm(ref_cfg, net_kwargs=dict(depth=spec["depth"]))
    R.load_det_weights(alg, seed=0, head_gain=spec["head_gain"])
    out = {}
    for it in range(STEPS):
        batch = {k: v.numpy() for k, v in batch_tensors(cfg, it).items()}
        alg.it = it
        b = {k: torch.from_numpy(v) for k, v in batch.items()}
        import inspect
        b = {k: v for k, v in b.items() if k in inspect.signature(alg.train_step).parameters}
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**b))
        loss = alg.out_dict["loss"]
        out[f"it{it}_loss"] = np.float32(loss.item())
        for k in ("train/sup_loss", "train/unsup_loss", "train/total_loss", "train/util_ratio"):
            out[f"it{it}_{k.split('/')[1]}"] = np.float32(alg.log_dict[k])
        alg.hooks_dict["ParamUpdateHook"].after_train_step(alg)
        sd = alg.model.state_dict()
        out[f"it{it}_head_bias"] = sd["head.bias"].numpy().copy()
        out[f"it{it}_qkv0_row0"] = sd["blocks.0.attn.qkv.weight"][0].numpy().copy()
        out[f"it{it}_param_sum"] = np.float64(sum(v.double().sum().item() for v in sd.values()))
        out[f"it{it}_rewarder_sum"] = np.float64(sum(v.double().sum().item() for v in alg.rewarder.state_dict().values()))
        h = alg.hooks_dict["MaskingHook"]
        if cfg["algorithm"] == "srflexmatch":
            out[f"it{it}_selected_label"] = h.selected_label.numpy().copy()
            out[f"it{it}_classwise_acc"] = h.classwise_acc.numpy().copy()
        elif cfg["algorithm"] == "srfreematch":
            out[f"it{it}_p_model"] = h.p_model.numpy().copy()
            out[f"it{it}_time_p"] = np.float32(h.time_p.item())
        elif cfg["algorithm"] == "srsoftmatch":
            out[f"it{it}_mu"] = np.float32(float(h.prob_max_mu_t))
            out[f"it{it}_var"] = np.float32(float(h.prob_max_var_t))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "->", len(out), "arrays")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    for name, spec in CASES.items():
        run_case(name, spec)
