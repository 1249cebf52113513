"""CPU, world_size 2, gloo: the host logic of the data-parallel path — gradient averaging of the flat
buffer, rank sharding, and that rank-local state (FlexMatch hook) is NOT synchronised."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from semireward_b200.parallel import allreduce_mean_, shard_indices
    flat = torch.full((1000,), float(rank + 1))
    allreduce_mean_(flat)
    ok_avg = bool(torch.allclose(flat, torch.full((1000,), (1 + world) / 2 * 1.0)))
    shard = shard_indices(10, rank, world)
    # weak scaling bookkeeping used by bench.py: value = world * samples * K / max over ranks of the time
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # rank-local FlexMatch state: oracle states fed different shards must stay different (never synchronised)
    from oracle import ssl_oracle as O
    st = O.FlexMatchState(16, 4)
    g = torch.Generator().manual_seed(rank)
    st.masking(torch.softmax(4 * torch.randn(4, 4, generator=g), -1), torch.arange(4) + 4 * rank, 0.5)
    # overlapped route of the ViT module (nets/vit.py): pieces of the flat gradient are all-reduced asynchronously as the
    # backward's block ranges finish, allreduce_grads_() completes them (gloo has no AVG: SUM, then divide)
    from semireward_b200.nets import vit_small_patch2_32
    net = vit_small_patch2_32(num_classes=10, depth=4)
    net._dp_group = dist.group.WORLD
    net._flat_grads = torch.full((100,), float(rank + 1))
    net._pending_reduce = [net._allreduce_async(net._flat_grads[60:], net._dp_group), net._allreduce_async(net._flat_grads[:60], net._dp_group)]
    net.allreduce_grads_()
    ok_overlap = bool(torch.allclose(net._flat_grads, torch.full((100,), (1 + world) / 2 * 1.0))) and net._pending_reduce == []
    bounds = [net._dp_bounds(12), net._dp_bounds(4), net._dp_bounds(2), net._dp_bounds(1)]
    q.put((rank, ok_avg and ok_overlap, shard, float(t.item()), st.selected_label.tolist(), bounds))
    dist.destroy_process_group()


def test_two_rank_gradient_average_and_sharding():
    world, port = 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert res[0][2] == [0, 2, 4, 6, 8] and res[1][2] == [1, 3, 5, 7, 9]
    assert res[0][3] == res[1][3] == 11.0
    assert res[0][4] != res[1][4]
    assert res[0][5] == [[9, 6, 3], [3, 2, 1], [1], []]   # dp_overlap_split = 4 block ranges


def _wrn_sync_worker(rank, world, port, q):
    """Data-parallel WRN: under DDP the reference's BatchNorms are SyncBatchNorms (misc.py:54), so the N
    ranks are NOT independent — every rank normalises with the statistics of all ranks' rows.  The oracle's sync form on
    `world` ranks must equal one process running BatchNorm over the union of the ranks' batches, with the mean of the
    per-rank losses as the loss (DDP averages gradients)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import torch.nn.functional as F
    from oracle import wrn_oracle as WO
    from semireward_b200 import detgen
    cfg = WO.WRNCfg(depth=10, num_classes=10)
    p = {n: torch.from_numpy(detgen.fill_param(n, s, 0)).requires_grad_(True) for n, s in cfg.param_shapes()}
    xs = [torch.from_numpy(detgen.normal("x", (6, 3, 16, 16), 100 + r)) for r in range(world)]
    ys = [torch.from_numpy(detgen.integers("y", (6,), 0, 10, 100 + r)) for r in range(world)]
    names = list(p)
    # this rank, synchronised statistics
    buf = WO.new_bn_buffers(cfg)
    logits, _ = WO.wrn_forward(p, buf, xs[rank], cfg, training=True, sync_group=dist.group.WORLD)
    loss = F.cross_entropy(logits, ys[rank])
    grads = torch.autograd.grad(loss, [p[n] for n in names], allow_unused=True)
    flat = torch.cat([(g if g is not None else torch.zeros_like(p[n])).reshape(-1) for n, g in zip(names, grads)])
    dist.all_reduce(flat)
    flat /= world                                   # DDP: average of the ranks' gradients
    # one process, BatchNorm over the union
    buf1 = WO.new_bn_buffers(cfg)
    lu, _ = WO.wrn_forward(p, buf1, torch.cat(xs), cfg, training=True)
    loss_u = sum(F.cross_entropy(lu[6 * r:6 * (r + 1)], ys[r]) for r in range(world)) / world
    gu = torch.autograd.grad(loss_u, [p[n] for n in names], allow_unused=True)
    flat_u = torch.cat([(g if g is not None else torch.zeros_like(p[n])).reshape(-1) for n, g in zip(names, gu)])
    err_l = (logits - lu[6 * rank:6 * (rank + 1)]).abs().max().item()
    err_g = (flat - flat_u).abs().max().item() / flat_u.abs().max().item()
    err_b = max((buf[k] - buf1[k]).abs().max().item() for k in buf)
    q.put((rank, err_l, err_g, err_b))
    dist.destroy_process_group()


def test_wrn_sync_batchnorm_equals_union_batch():
    world, port = 2, 29741
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_wrn_sync_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err_l, err_g, err_b in res:
        assert err_l < 2e-4 and err_g < 2e-4 and err_b < 1e-5, (rank, err_l, err_g, err_b)


def _oracle_dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import batch_tensors, build_oracle, small_cfg
    cfg = small_cfg(start_timing=2, N_k=2, num_train_iter=16)
    orc = build_oracle(cfg, 1)
    orc.dp_group = dist.group.WORLD
    for it in range(3):
        orc.train_step(dict(batch_tensors(cfg, it, seed=1 + rank)), it)
        orc.param_update()
    q.put((rank, {k: v.detach().numpy().copy() for k, v in orc.p.items()}, {k: v.detach().numpy().copy() for k, v in orc.rp.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_oracle_data_parallel_semantics_two_ranks():
    """The oracle's N-rank form (the checker of the 2-GPU NCCL parity test): backbone gradients averaged -> the ranks' backbones stay
    bit-identical; of the SR update only the generator-loss gradient is averaged (DDP synchronises the first backward only,
    scripts/c3_ddp_probe.py) -> the ranks' Rewarders differ, as they do in the reference."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_oracle_dp_worker, args=(r, 2, 29641, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        r, p, rp = q.get(timeout=600)
        got[r] = (p, rp)
    for p in procs:
        p.join(60)
    import numpy as np
    for k in got[0][0]:
        assert np.array_equal(got[0][0][k], got[1][0][k]), k
    assert any(not np.array_equal(got[0][1][k], got[1][1][k]) for k in got[0][1])
