"""CPU, world_size 2, gloo: the host logic of the data-parallel path — gradient averaging of the flat
buffer, rank sharding, and that rank-local state (FlexMatch hook) is NOT synchronised."""
import os

import torch
import torch.distributed as dist
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
