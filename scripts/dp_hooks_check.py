"""torchrun check (2+ ranks): SRFreeMatch / SRSoftMatch under data parallelism (NCCL all_gather of the probabilities / max
probabilities).  The hook statistics integrate every rank's rows, so after a few steps on DIFFERENT per-rank
batches the hook state must be bit-identical across ranks, while the masks differ.
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_hooks_check.py"""
import faulthandler
import functools
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(150, exit=True)
import bench  # noqa: E402
import semireward_b200 as S  # noqa: E402
from semireward_b200 import detgen  # noqa: E402
from semireward_b200.parallel import send_model_cuda  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok_all = True
for algorithm in ("srfreematch", "srsoftmatch"):
    cfg = dict(bench.YAML_CFG, algorithm=algorithm, gpu=local, distributed=True, world_size=world, rank=rank, num_train_iter=64, start_timing=2,
               N_k=2, ema_p=0.9, use_quantile=True, clip_thresh=False, ent_loss_ratio=0.01, hard_label=True, T=0.5, dist_align=True,
               dist_uniform=True, n_sigma=2, per_class=False)
    args = S.get_config(cfg)
    torch.manual_seed(0)
    alg = S.get_algorithm(args, functools.partial(S.get_net_builder(args.net, False), depth=2, drop_path_rate=0.1), None, None)
    alg.model = send_model_cuda(args, alg.model)
    alg.model.train()
    masks = []
    for it in range(5):   # stage 1, the gap step, stage 2 (batched stochastic passes)
        b = detgen.ssl_batch(8, 1, 100, 50000, seed=1 + rank, step=it)
        alg.it = it
        alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**{k: torch.from_numpy(v) for k, v in b.items()}))
        alg.call_hook("after_train_step", "ParamUpdateHook")
        masks.append(alg._last_mask.clone())
    torch.cuda.synchronize()
    h = alg.hooks_dict["MaskingHook"]
    if algorithm == "srfreematch":
        state = torch.cat([h.time_p.flatten(), h.p_model.flatten(), h.label_hist.flatten()])
    else:
        state = torch.cat([h.prob_max_mu_t.flatten(), h.prob_max_var_t.flatten(), alg.hooks_dict["DistAlignHook"].p_model.flatten()])
    params = torch.cat([p.detach().flatten() for p in alg._net().parameters()])
    g_state = [torch.empty_like(state) for _ in range(world)]
    g_par = [torch.empty_like(params) for _ in range(world)]
    dist.all_gather(g_state, state)
    dist.all_gather(g_par, params)
    same_state = all(torch.equal(g_state[0], g) for g in g_state)
    same_par = all(torch.equal(g_par[0], g) for g in g_par)
    finite = bool(torch.isfinite(state).all()) and bool(torch.isfinite(params).all())
    print(f"rank {rank} {algorithm}: hook state identical across ranks {same_state}, parameters identical {same_par}, finite {finite}, "
          f"total_loss {alg.log_dict['train/total_loss']:.4f}", flush=True)
    ok_all = ok_all and same_state and same_par and finite
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
