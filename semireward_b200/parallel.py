"""Data parallelism — the only strategy the reference has: one process per GPU, every rank runs the
full train_step on its own batch, backbone (and Rewarder) gradients are averaged over the ranks with NCCL over
NVLink/NVSwitch.  Mirrors send_model_cuda (semilearn/core/utils/misc.py:39-70) and the DDP wrapper's `.module` surface.

The native backward produces ALL gradients of the step as one flat fp32 buffer, so the exchange is a single
all-reduce(avg) of that buffer (85.7 MB for ViT-S) instead of DDP's bucket hooks.  FlexMatch hook state, the Rewarder's
batch context and the reward-mean threshold stay rank-local exactly as in the reference."""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place average over the ranks of `group`."""
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:  # gloo (CPU tests) has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    return flat


class NativeDataParallel(nn.Module):
    """DDP-shaped wrapper (`.module`, forwards every call) that tells the native autograd function to average the flat
    gradient buffer over `process_group` at the end of its backward."""

    def __init__(self, module: nn.Module, process_group=None):
        super().__init__()
        self.module = module
        self.process_group = process_group if process_group is not None else dist.group.WORLD
        module._dp_group = self.process_group

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


def send_model_cuda(args, model, clip_batch=True):
    if not torch.cuda.is_available():
        raise Exception("ONLY GPU TRAINING IS SUPPORTED")
    if getattr(args, "distributed", False):
        gpu = args.gpu if args.gpu is not None else torch.cuda.current_device()
        torch.cuda.set_device(gpu)
        model = model.cuda(gpu)
        # NB the reference divides args.batch_size by ngpus here *after* the loaders were built with the YAML value
        # (SURVEY.md §8e): the YAML batch is per GPU, so nothing is rescaled on this path.
        with torch.no_grad():
            for p in model.parameters():  # DDP broadcasts rank 0's parameters at construction
                dist.broadcast(p.data, src=0)
        return NativeDataParallel(model)
    return model.cuda(args.gpu)


def shard_indices(n_items: int, rank: int, world_size: int):
    """rank r takes items r::W (DistributedSampler stride, semilearn/datasets/samplers/sampler.py:70)."""
    return list(range(rank, n_items, world_size))
