"""C3: what does torch DDP do with the reference's SR update, which calls backward TWICE after ONE forward of
the DDP-wrapped Rewarder (srflexmatch.py:204-205: generator_loss.backward(retain_graph=True); rewarder_loss.backward(...))?

Runs the reference's own Rewarder class under DistributedDataParallel(find_unused_parameters=True, broadcast_buffers=False)
(misc.py:56-58) on 2 gloo ranks (CPU, build container only) with different features per rank and prints, per rank, how the
gradient that reaches the optimizer relates to the two local gradients g1 (generator_loss) and g2 (rewarder_loss):
    both synchronised  -> grad == mean_r(g1 + g2)          (identical on both ranks)
    first only         -> grad == mean_r(g1) + local g2    (differs between ranks -> the ranks' Rewarders drift apart)
Usage: python scripts/c3_ddp_probe.py        (writes nothing; the finding is recorded in DESIGN.md §7)"""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ref_driver as R
    R.load_reference()
    from semilearn.algorithms.semireward.semireward import Rewarder
    torch.manual_seed(0)
    net = Rewarder(100, 128, 384)
    ddp = torch.nn.parallel.DistributedDataParallel(net, broadcast_buffers=False, find_unused_parameters=True)
    g = torch.Generator().manual_seed(100 + rank)
    feats = torch.randn(8, 384, generator=g)
    gen = torch.randint(0, 100, (8,), generator=g)
    target = torch.where(torch.rand(8, 1, generator=g) > 0.5, 1.0, 0.5)
    crit = torch.nn.MSELoss()
    # local gradients without DDP
    r0 = net(feats, gen)
    names = [n for n, _ in net.named_parameters()]
    g1 = torch.autograd.grad(crit(r0, torch.ones_like(r0)), list(net.parameters()), retain_graph=True, allow_unused=True)
    g2 = torch.autograd.grad(crit(r0, target), list(net.parameters()), allow_unused=True)
    z = lambda t, p: torch.zeros_like(p) if t is None else t   # noqa: E731
    g1 = [z(a, p) for a, p in zip(g1, net.parameters())]
    g2 = [z(a, p) for a, p in zip(g2, net.parameters())]
    # the reference's sequence
    net.zero_grad()
    reward = ddp(feats, gen)
    crit(reward, torch.ones_like(reward)).backward(retain_graph=True)
    after_first = [None if p.grad is None else p.grad.clone() for p in net.parameters()]
    crit(reward, target).backward(retain_graph=True)
    got = [torch.zeros_like(p) if p.grad is None else p.grad.clone() for p in net.parameters()]

    def allmean(ts):
        out = []
        for t in ts:
            t = t.clone()
            dist.all_reduce(t)
            out.append(t / world)
        return out
    m1, m2 = allmean(g1), allmean(g2)
    both = max((a - (b + c)).abs().max().item() for a, b, c in zip(got, m1, m2))
    first_only = max((a - (b + c)).abs().max().item() for a, b, c in zip(got, m1, g2))
    none = max((a - (b + c)).abs().max().item() for a, b, c in zip(got, g1, g2))
    scale = max(a.abs().max().item() for a in got)
    print(f"rank {rank}: |grad - mean(g1+g2)| = {both:.3e}   |grad - (mean g1 + local g2)| = {first_only:.3e}   "
          f"|grad - local(g1+g2)| = {none:.3e}   (grad scale {scale:.3e})", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    mp.spawn(worker, args=(2, 29617), nprocs=2, join=True)
