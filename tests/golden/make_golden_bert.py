"""Generate the text-path golden vectors from the LIVE reference (build container only; /root/reference must be mounted).

    python tests/golden/make_golden_bert.py

Drives the reference's own SRSoftMatch / SRFixMatch train_step + ParamUpdateHook on CPU with `net: bert_base_uncased`,
`use_cat: False` through oracle/ref_driver.py: `BertModel.from_pretrained` (no hub access here) is handed a randomly
initialised 2-layer `BertModel(BertConfig(...))` of transformers 5.5.0 with eager attention and every dropout at 0, the
weights are overwritten with semireward_b200.detgen fills and the batches come from detgen.nlp_batch (padding tails
included).  oracle/bert_oracle.py is pinned against these files by tests/test_bert_oracle.py on any machine."""
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

from golden_cases import BERT_CASES, BERT_SMALL, STEPS, bert_small_cfg  # noqa: E402


def text_batch(cfg, it):
    from semireward_b200 import detgen
    b = detgen.nlp_batch(cfg["batch_size"], cfg["uratio"], cfg["num_classes"], cfg["ulb_dest_len"], max_length=BERT_SMALL["max_length"],
                         vocab_size=BERT_SMALL["vocab_size"], seed=1, step=it)
    return {k: ({kk: torch.from_numpy(vv) for kk, vv in v.items()} if isinstance(v, dict) else torch.from_numpy(v)) for k, v in b.items()}


def run_case(name, spec, attn="eager"):
    from oracle import ref_driver as R
    cfg = bert_small_cfg(**spec["cfg"])
    ref_cfg = dict(cfg)
    ref_cfg.update(ema_p=0.999, ent_loss_ratio=0.001, use_quantile=True, clip_thresh=False, dist_align=True, dist_uniform=True, n_sigma=2,
                   per_class=False)
    hf = dict(vocab_size=BERT_SMALL["vocab_size"], num_hidden_layers=BERT_SMALL["layers"], max_position_embeddings=BERT_SMALL["max_position"],
              hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, attn_implementation=attn)
    alg = R.build_reference_algorithm(ref_cfg, net_kwargs=dict(bert=hf, dropout=0.0))
    R.load_det_weights(alg, seed=0, head_gain=spec["head_gain"])
    out = {}
    for it in range(STEPS):
        b = text_batch(cfg, it)
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
        out[f"it{it}_q0_row0"] = sd["bert.encoder.layer.0.attention.self.query.weight"][0].numpy().copy()
        out[f"it{it}_word_row7"] = sd["bert.embeddings.word_embeddings.weight"][7].numpy().copy()
        out[f"it{it}_rewarder_sum"] = np.float64(sum(v.double().sum().item() for v in alg.rewarder.state_dict().values()))
        if cfg["algorithm"] == "srsoftmatch":
            h = alg.hooks_dict["MaskingHook"]
            out[f"it{it}_mu"] = np.float32(float(h.prob_max_mu_t))
            out[f"it{it}_var"] = np.float32(float(h.prob_max_var_t))
    return out


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    for name, spec in BERT_CASES.items():
        out = run_case(name, spec)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "->", len(out), "arrays")
