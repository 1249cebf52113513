"""Generate the WRN golden vectors from the LIVE reference (build container only; /root/reference must be mounted).

    python tests/golden/make_golden_wrn.py

Drives the reference's own SRFlexMatch / SRFixMatch train_step + ParamUpdateHook on CPU with a WideResNet (depth 10, widen
2: the class behind `net: wrn_28_2`), `use_cat: True`, SGD(momentum 0.9, nesterov) — the optimizer and net of BASELINE
configs[0] — through oracle/ref_driver.py, with semireward_b200.detgen weights and batches.  oracle/wrn_oracle.py is pinned
against these files by tests/test_wrn_oracle.py on any machine."""
from __future__ import annotations

import inspect
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden_cases import STEPS, WRN_CASES, wrn_small_cfg  # noqa: E402


def image_batch(cfg, it):
    from semireward_b200 import detgen
    b = detgen.ssl_batch(cfg["batch_size"], cfg["uratio"], cfg["num_classes"], cfg["ulb_dest_len"], seed=1, step=it)
    return {k: torch.from_numpy(v) for k, v in b.items()}


def run_case(name, spec):
    from oracle import ref_driver as R
    cfg = wrn_small_cfg(**spec["cfg"])
    alg = R.build_reference_algorithm(dict(cfg), net_kwargs=dict(depth=spec["depth"]))
    R.load_det_weights(alg, seed=0, head_gain=spec["head_gain"])
    out = {}
    for it in range(STEPS):
        b = image_batch(cfg, it)
        alg.it = it
        b = {k: v for k, v in b.items() if k in inspect.signature(alg.train_step).parameters}
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**b))
        out[f"it{it}_loss"] = np.float32(alg.out_dict["loss"].item())
        for k in ("train/sup_loss", "train/unsup_loss", "train/util_ratio"):
            out[f"it{it}_{k.split('/')[1]}"] = np.float32(alg.log_dict[k])
        alg.hooks_dict["ParamUpdateHook"].after_train_step(alg)
        sd = alg.model.state_dict()
        out[f"it{it}_cls_bias"] = sd["classifier.bias"].numpy().copy()
        out[f"it{it}_conv1_w0"] = sd["conv1.weight"][0].numpy().copy()
        out[f"it{it}_b3_bn1_mean"] = sd["block3.layer.0.bn1.running_mean"].numpy().copy()     # advanced although its output is unused
        out[f"it{it}_b3_bn1_weight"] = sd["block3.layer.0.bn1.weight"].numpy().copy()         # never updated (no gradient)
        out[f"it{it}_bn1_var"] = sd["bn1.running_var"].numpy().copy()
        out[f"it{it}_param_sum"] = np.float64(sum(v.double().sum().item() for n, v in sd.items() if not n.endswith("num_batches_tracked")))
        out[f"it{it}_rewarder_sum"] = np.float64(sum(v.double().sum().item() for v in alg.rewarder.state_dict().values()))
        if cfg["algorithm"] == "srflexmatch":
            h = alg.hooks_dict["MaskingHook"]
            out[f"it{it}_selected_label"] = h.selected_label.numpy().copy()
            out[f"it{it}_classwise_acc"] = h.classwise_acc.numpy().copy()
    return out


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    for name, spec in WRN_CASES.items():
        out = run_case(name, spec)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "->", len(out), "arrays; util", [float(out[f"it{i}_util_ratio"]) for i in range(STEPS)])
