"""Dict -> argparse.Namespace config, mirroring semilearn/lighting/config.py:11-153 (`get_config(dict)`): the keys are
the reference's YAML keys, unknown keys are attached verbatim (over_write_args_from_dict behaviour, misc.py:18-27),
algorithm-specific defaults come from `Algorithm.get_argument()` (train.py:248-254)."""
from __future__ import annotations

import argparse

_DEFAULTS = dict(
    save_dir="./saved_models", save_name="srflexmatch", resume=False, load_path=None, overwrite=True, use_tensorboard=False,
    use_wandb=False, use_aim=False, epoch=1, num_train_iter=20, num_warmup_iter=0, num_eval_iter=10, num_log_iter=5,
    num_labels=400, batch_size=8, uratio=1, eval_batch_size=16, ema_m=0.999, ulb_loss_ratio=1.0, optim="SGD", lr=3e-2,
    momentum=0.9, weight_decay=5e-4, layer_decay=1.0, net="vit_small_patch2_32", net_from_name=False, use_pretrain=False,
    pretrain_path="", algorithm="srflexmatch", use_cat=True, amp=False, clip_grad=0, imb_algorithm=None, data_dir="./data",
    dataset="synthetic", num_classes=100, train_sampler="RandomSampler", num_workers=1, include_lb_to_ulb=True, lb_imb_ratio=1,
    ulb_imb_ratio=1, ulb_num_labels=None, img_size=32, crop_ratio=0.875, max_length=512, max_length_seconds=4.0, sample_rate=16000,
    seed=1, world_size=1, rank=0, dist_url="tcp://127.0.0.1:10001", dist_backend="nccl", gpu=0, multiprocessing_distributed=False,
    distributed=False,
)


def get_config(config: dict) -> argparse.Namespace:
    from . import ALGORITHMS  # late import: algorithms register themselves on package import
    args = argparse.Namespace(**_DEFAULTS)
    alg = config.get("algorithm", args.algorithm)
    if alg not in ALGORITHMS:
        raise KeyError(f"Unknown algorithm: {alg}")
    for a in ALGORITHMS[alg].get_argument():
        setattr(args, a.name.lstrip("-"), a.default)
    for k, v in config.items():
        setattr(args, k, v)
    return args
