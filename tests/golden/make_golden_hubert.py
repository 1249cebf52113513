"""Generate the audio-path golden vectors from the LIVE reference (build container only; /root/reference must be mounted).

    python tests/golden/make_golden_hubert.py

Drives the reference's own SRFlexMatch / SRFixMatch train_step + ParamUpdateHook on CPU with `net: hubert_base`,
`use_cat: False` through oracle/ref_driver.py: `HubertModel.from_pretrained` (no hub access here) is handed a randomly
initialised 2-layer `HubertModel(HubertConfig(...))` of transformers 5.5.0 with the full convolutional stem, eager attention
and every source of randomness off (dropouts, LayerDrop, SpecAugment); weights and clips come from semireward_b200.detgen.
oracle/hubert_oracle.py is pinned against these files by tests/test_hubert_oracle.py on any machine."""
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

from golden_cases import HUBERT_CASES, HUBERT_SMALL, STEPS, hubert_small_cfg  # noqa: E402


def audio_batch(cfg, it):
    from semireward_b200 import detgen
    b = detgen.audio_batch(cfg["batch_size"], cfg["uratio"], cfg["num_classes"], cfg["ulb_dest_len"], samples=HUBERT_SMALL["samples"], seed=1, step=it)
    return {k: torch.from_numpy(v) for k, v in b.items()}


def run_case(name, spec, attn="eager"):
    from oracle import ref_driver as R
    cfg = hubert_small_cfg(**spec["cfg"])
    hf = dict(num_hidden_layers=HUBERT_SMALL["layers"], hidden_dropout=0.0, activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0,
              final_dropout=0.0, layerdrop=0.0, apply_spec_augment=False, attn_implementation=attn)
    alg = R.build_reference_algorithm(dict(cfg), net_kwargs=dict(hubert=hf, dropout=0.0))
    R.load_det_weights(alg, seed=0, head_gain=spec["head_gain"])
    out = {}
    for it in range(STEPS):
        b = audio_batch(cfg, it)
        alg.it = it
        b = {k: v for k, v in b.items() if k in inspect.signature(alg.train_step).parameters}
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**b))
        out[f"it{it}_loss"] = np.float32(alg.out_dict["loss"].item())
        for k in ("train/sup_loss", "train/unsup_loss", "train/util_ratio"):
            out[f"it{it}_{k.split('/')[1]}"] = np.float32(alg.log_dict[k])
        out[f"it{it}_feat_lb"] = alg.out_dict["feat"]["x_lb"].detach().numpy().copy()
        alg.hooks_dict["ParamUpdateHook"].after_train_step(alg)
        sd = alg.model.state_dict()
        out[f"it{it}_cls_bias"] = sd["classifier.2.bias"].numpy().copy()
        out[f"it{it}_conv0_w0"] = sd["model.feature_extractor.conv_layers.0.conv.weight"][0, 0].numpy().copy()
        out[f"it{it}_pos_g"] = sd["model.encoder.pos_conv_embed.conv.parametrizations.weight.original0"].reshape(-1).numpy().copy()
        out[f"it{it}_q0_row0"] = sd["model.encoder.layers.0.attention.q_proj.weight"][0].numpy().copy()
        out[f"it{it}_rewarder_sum"] = np.float64(sum(v.double().sum().item() for v in alg.rewarder.state_dict().values()))
        if cfg["algorithm"] == "srflexmatch":
            h = alg.hooks_dict["MaskingHook"]
            out[f"it{it}_selected_label"] = h.selected_label.numpy().copy()
            out[f"it{it}_classwise_acc"] = h.classwise_acc.numpy().copy()
    return out


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    for name, spec in HUBERT_CASES.items():
        out = run_case(name, spec)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "->", len(out), "arrays; util", [float(out[f"it{i}_util_ratio"]) for i in range(STEPS)])
