"""Minimal NCCL sanity check under torchrun: broadcast + all_reduce(avg) of an 85.7 MB buffer, timed on the device."""
import faulthandler
import os
import time

import torch
import torch.distributed as dist

faulthandler.dump_traceback_later(60, exit=True)
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
t0 = time.time()
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.full((21_400_000,), float(rank + 1), device="cuda")
dist.broadcast(x, src=0)
torch.cuda.synchronize()
print(f"rank {rank}: init+broadcast {time.time() - t0:.1f} s", flush=True)
x.fill_(float(rank + 1))
for _ in range(3):
    dist.all_reduce(x, op=dist.ReduceOp.AVG)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    dist.all_reduce(x, op=dist.ReduceOp.AVG)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"rank {rank}: all_reduce(avg) of {x.numel() * 4 / 1e6:.1f} MB: {ms:.3f} ms -> bus {2 * (world - 1) / world * x.numel() * 4 / ms / 1e6:.0f} GB/s, value {x[0].item():.3f}", flush=True)
dist.barrier()
dist.destroy_process_group()
