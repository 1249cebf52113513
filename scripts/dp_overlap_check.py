"""torchrun check (2+ ranks): the overlapped gradient all-reduce (backward split in two block ranges, tail of the flat buffer
reduced while the rest of the backward runs) gives bit-identical averaged gradients to one all-reduce after the backward.
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_overlap_check.py"""
import faulthandler
import functools
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(120, exit=True)
import bench  # noqa: E402
import semireward_b200 as S  # noqa: E402
from semireward_b200 import detgen  # noqa: E402
from semireward_b200.parallel import send_model_cuda  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
flats = []
for split in (0, 2, 0, 4):
    cfg = dict(bench.YAML_CFG, gpu=local, distributed=True, world_size=world, rank=rank, num_train_iter=64, start_timing=1000)
    args = S.get_config(cfg)
    torch.manual_seed(0)
    alg = S.get_algorithm(args, functools.partial(S.get_net_builder(args.net, False), depth=4, drop_path_rate=0.0), None, None)
    alg.model = send_model_cuda(args, alg.model)
    alg.model.train()
    alg._net().dp_overlap_split = split
    for it in (1, 2, 3):   # 3 steps: eager, captured, replayed
        b = detgen.ssl_batch(8, 1, 100, 50000, seed=1 + rank, step=it)
        alg.it = it
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**{k: torch.from_numpy(v) for k, v in b.items()}))
        flat = alg._net()._flat_grads
        torch.cuda.synchronize()
        alg.call_hook("after_train_step", "ParamUpdateHook")
    torch.cuda.synchronize()
    flats.append(torch.cat([p.detach().flatten() for p in alg._net().parameters()]).clone())
ok = all(torch.equal(flats[0], f) for f in flats[1:])
gathered = [torch.empty_like(flats[0]) for _ in range(world)]
dist.all_gather(gathered, flats[1])
same_across_ranks = all(torch.equal(gathered[0], g) for g in gathered)
print(f"rank {rank}: parameters after 3 steps identical for 1 / 2 / 1 / 4 block ranges: {ok}; identical across ranks: {same_across_ranks}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok and same_across_ranks else 1)
