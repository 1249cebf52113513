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
    q.put((rank, ok_avg, shard, float(t.item()), st.selected_label.tolist()))
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
