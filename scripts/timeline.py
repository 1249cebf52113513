"""GPU timeline of the bench step from CUPTI (torch.profiler): per-kernel busy time, idle gaps and where they are.
python scripts/timeline.py [--steps 6] [--batch-size 8] [--out gpurun_out/timeline.txt]
Numbers taken under the profiler are diagnostic only (never bench values)."""
import argparse
import collections
import json
import os
import re
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import semireward_b200 as S  # noqa: E402
from semireward_b200 import detgen  # noqa: E402
from semireward_b200.parallel import send_model_cuda  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--batch-size", type=int, default=8)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    cfg = dict(bench.YAML_CFG, batch_size=a.batch_size, gpu=0, distributed=False, world_size=1, rank=0)
    args = S.get_config(cfg)
    torch.manual_seed(0)
    alg = S.get_algorithm(args, S.get_net_builder(args.net, False), None, None)
    alg.model = send_model_cuda(args, alg.model)
    alg.model.train()
    B = args.batch_size
    batches = []
    for i in range(4):
        b = detgen.ssl_batch(B, args.uratio, args.num_classes, args.ulb_dest_len, seed=1, step=i)
        batches.append({k: torch.from_numpy(v).cuda() for k, v in b.items()})

    def step(i):
        alg.it = 1 + i
        alg.out_dict, alg.log_dict = alg.train_step(**batches[i % 4])
        alg.call_hook("after_train_step", "ParamUpdateHook")   # the metric excludes the EMA / logging hooks

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(a.steps):
            step(5 + i)
        torch.cuda.synchronize()
    path = os.path.join(tempfile.gettempdir(), "srw_trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    lines = []
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    span = t1 - t0
    # union of busy intervals (streams overlap)
    busy, cur_s, cur_e = 0.0, None, None
    gaps = []
    for e in ev:
        s, f = e["ts"], e["ts"] + e["dur"]
        if cur_e is None:
            cur_s, cur_e = s, f
        elif s <= cur_e:
            cur_e = max(cur_e, f)
        else:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, e["name"]))
            cur_s, cur_e = s, f
    busy += cur_e - cur_s
    lines.append(f"steps {a.steps} span {span / a.steps:.1f} us/step, GPU busy (union) {busy / a.steps:.1f} us/step, idle {100 * (1 - busy / span):.1f}%, "
                 f"{len(ev) / a.steps:.0f} GPU activities/step")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in ev:
        n = re.sub(r"\(.*", "", e["name"])[:70]
        agg[n][0] += 1
        agg[n][1] += e["dur"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        lines.append(f"{k:70s} n/step={n / a.steps:6.1f} us/step={t / a.steps:8.1f} avg={t / n:7.1f}")
    gagg = collections.defaultdict(lambda: [0, 0.0])
    for g, n in gaps:
        n = re.sub(r"\(.*", "", n)[:70]
        gagg[n][0] += 1
        gagg[n][1] += g
    lines.append("-- idle time BEFORE kernel (sum over run / per step) --")
    for k, (n, t) in sorted(gagg.items(), key=lambda kv: -kv[1][1])[:25]:
        lines.append(f"{k:70s} n/step={n / a.steps:6.1f} idle us/step={t / a.steps:8.1f} avg={t / n:6.2f}")
    big = sorted(gaps, key=lambda g: -g[0])[:12]
    lines.append("-- largest single gaps --")
    for g, n in big:
        lines.append(f"{g:8.1f} us before {n[:80]}")
    txt = "\n".join(lines)
    print(txt)
    if a.out:
        open(a.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
