"""-m gpu, needs >= 2 devices (skipped otherwise): N-rank parity under NCCL.  Every rank runs the native
SRFlexMatch step on ITS OWN shard with the backbone wrapped by send_model_cuda and, in the same process, the oracle on the same
shard with its gradients averaged over a gloo group ("N independent single-rank oracles + averaged gradients").  Checked per
rank and step: pseudo-labels / masks / FlexMatch hook state bit-exact (rank-local, never synchronised), losses 1e-3, the averaged
backbone gradient 1e-3 relative, Rewarder parameters after the DDP-style update (generator-loss gradient averaged,
rewarder-loss gradient local — scripts/c3_ddp_probe.py); across ranks: backbones bit-identical, hook state different."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        torch.cuda.set_device(rank)
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        sys.path.insert(0, ROOT)
        from helpers import batch_tensors, build_native, build_oracle, small_cfg
        from test_train_step_gpu import _grad_tap, _resync
        from semireward_b200.parallel import send_model_cuda
        torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
        cfg = small_cfg(gpu=rank, distributed=True, world_size=world, rank=rank, start_timing=2, N_k=2, num_train_iter=16)
        orc = build_oracle(cfg, 2)
        orc.dp_group = dist.group.WORLD
        alg = build_native(cfg, 2)
        alg.model = send_model_cuda(alg.args, alg.model)
        alg.rewarder._dp_group = dist.group.WORLD
        hook = alg.hooks_dict["MaskingHook"]
        tap = _grad_tap(alg)
        log = []
        for it in range(5):
            batch = batch_tensors(cfg, it, seed=1 + rank)      # a different shard per rank
            rec = orc.train_step(dict(batch), it)
            ref_grads = orc.param_update()
            alg.it = it
            alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**batch))
            alg.call_hook("after_train_step")
            torch.cuda.synchronize()
            ld = alg.log_dict
            assert torch.equal(alg._last_pseudo_label.cpu(), rec["pseudo"]) and torch.equal(alg._last_mask.cpu(), rec["mask"]), f"rank {rank} it {it}"
            assert torch.equal(hook.selected_label.cpu(), orc.hook.selected_label) and torch.equal(hook.classwise_acc.cpu(), orc.hook.classwise_acc)
            for kn, ko in (("train/sup_loss", "sup_loss"), ("train/total_loss", "total_loss")):
                assert abs(ld[kn] - float(rec[ko])) < 1e-3, f"rank {rank} it {it} {ko}: {ld[kn]} vs {float(rec[ko])}"
            worst = 0.0
            for n, p in alg._net().named_parameters():
                gr = ref_grads[n]
                worst = max(worst, (tap[n].cpu() - gr).abs().max().item() / max(gr.abs().max().item(), 1e-20))
            assert worst < 1e-3, f"rank {rank} it {it}: averaged gradient error {worst}"
            worst_r = 0.0
            for n, p in alg.rewarder.named_parameters():
                dp = (p.detach().cpu() - orc.rp[n].detach()).abs()
                gr = rec.get("sr_grads", {}).get(n)
                if gr is not None:
                    well = gr.abs() > max(1e-2 * gr.abs().max().item(), 1e-6)
                    dp = dp[well] if well.any() else dp[:0]
                if dp.numel():
                    worst_r = max(worst_r, dp.max().item())
            assert worst_r < 2e-5, f"rank {rank} it {it}: rewarder parameters {worst_r}"
            log.append(f"rank {rank} it {it}: total {ld['train/total_loss']:.5f} (oracle {float(rec['total_loss']):.5f}) avg-grad rel err {worst:.2e} rewarder {worst_r:.2e}")
            # resync from the oracle (itself identical on every rank for the backbone)
            with torch.no_grad():
                for n, p in alg._net().named_parameters():
                    p.copy_(orc.p[n].detach())
                for n, p in alg.rewarder.named_parameters():
                    p.copy_(orc.rp[n].detach())
        # free-running: two more steps without resync must keep the ranks' backbones bit-identical
        for it in (5, 6):
            alg.it = it
            alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**batch_tensors(cfg, it, seed=1 + rank)))
            alg.call_hook("after_train_step")
        torch.cuda.synchronize()
        flat = torch.cat([p.detach().flatten() for p in alg._net().parameters()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        sel = hook.selected_label.clone()
        sels = [torch.empty_like(sel) for _ in range(world)]
        dist.all_gather(sels, sel)
        differ = any(not torch.equal(sels[0], s) for s in sels[1:])
        q.put((rank, "ok", same, differ, log))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:   # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc(), None, []))
        raise


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_two_rank_nccl_step_parity_vs_rank_oracles():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29653, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(120)
    for rank, status, same, differ, log in sorted(res):
        assert status == "ok", same
        print("\n".join(log))
        assert same, "backbone parameters differ between ranks after NCCL-averaged steps"
        assert differ, "FlexMatch selected_label must stay rank-local (different shards -> different state)"


def _wrn_worker(rank, world, port, q):
    """WRN under data parallel = SyncBatchNorm (misc.py:54) + gradient averaging: statistics over BOTH ranks' rows in every BatchNorm,
    forward and backward, against the oracle's SyncBatchNorm form (differentiable gloo all-reduces) on the same shards."""
    import functools
    import torch.distributed as dist
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        torch.cuda.set_device(rank)
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        sys.path.insert(0, ROOT)
        import semireward_b200 as S
        from golden_cases import wrn_small_cfg
        from oracle import ssl_oracle as O, wrn_oracle as WO
        from semireward_b200 import detgen
        from semireward_b200.nets.wrn import WideResNet
        from semireward_b200.parallel import send_model_cuda
        from test_train_step_gpu import _grad_tap, _mask2_tie
        torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
        cfg = wrn_small_cfg(gpu=rank, distributed=True, world_size=world, rank=rank, start_timing=2, N_k=2, num_train_iter=16)
        sc = O.StepConfig(algorithm=cfg["algorithm"], num_classes=cfg["num_classes"], ulb_dest_len=cfg["ulb_dest_len"], p_cutoff=cfg["p_cutoff"],
                          thresh_warmup=cfg["thresh_warmup"], start_timing=cfg["start_timing"], N_k=cfg["N_k"], num_train_iter=cfg["num_train_iter"],
                          num_warmup_iter=cfg["num_warmup_iter"], lr=cfg["lr"], weight_decay=cfg["weight_decay"], layer_decay=cfg["layer_decay"],
                          sr_lr=cfg["sr_lr"], feature_dim=cfg["feature_dim"])
        wc = WO.WRNCfg(depth=10, num_classes=cfg["num_classes"])
        orc = WO.build_det_wrn_oracle(wc, sc, seed=0, head_gain=4.0)
        orc.dp_group = dist.group.WORLD
        args = S.get_config(cfg)
        alg = S.get_algorithm(args, functools.partial(WideResNet, first_stride=1, depth=10, widen_factor=2), None, None)
        with torch.no_grad():
            for prefix, mod in (("", alg.model), ("rewarder.", alg.rewarder), ("generator.", alg.generator)):
                for n, p in mod.named_parameters():
                    p.copy_(torch.from_numpy(detgen.fill_param(prefix + n, p.shape, 0)))
                    if prefix == "" and n == "classifier.weight":
                        p.mul_(4.0)
        alg.model = alg.model.cuda(rank).train()
        alg.rewarder, alg.generator = alg.rewarder.cuda(rank), alg.generator.cuda(rank)
        alg.model = send_model_cuda(alg.args, alg.model)
        alg.rewarder._dp_group = dist.group.WORLD
        tap = _grad_tap(alg)
        log = []
        for it in range(5):
            b = detgen.ssl_batch(cfg["batch_size"], cfg["uratio"], cfg["num_classes"], cfg["ulb_dest_len"], seed=1 + rank, step=it)   # a different shard per rank
            batch = {k: torch.from_numpy(v) for k, v in b.items()}
            rec = orc.train_step(dict(batch), it)
            ref_grads = orc.param_update()
            alg.it = it
            alg.out_dict, alg.log_dict = alg.train_step(**alg.process_batch(**batch))
            alg.call_hook("after_train_step")
            torch.cuda.synchronize()
            ld = alg.log_dict
            assert torch.equal(alg._last_pseudo_label.cpu(), rec["pseudo"]) and torch.equal(alg._last_mask.cpu(), rec["mask"]), f"rank {rank} it {it}"
            assert abs(ld["train/sup_loss"] - float(rec["sup_loss"])) < 1e-3, (rank, it, ld, float(rec["sup_loss"]))
            worst_b = max((bf.cpu() - orc.buf[n]).abs().max().item() for n, bf in alg._net().named_buffers() if not n.endswith("num_batches_tracked"))
            assert worst_b < 1e-5, (rank, it, worst_b)
            if not _mask2_tie(rec):
                assert abs(ld["train/total_loss"] - float(rec["total_loss"])) < 1e-3, (rank, it, ld, float(rec["total_loss"]))
                worst_l2 = 0.0
                gmax = max(g.abs().max().item() for g in ref_grads.values() if g is not None)
                for n, p in alg._net().named_parameters():
                    gr = ref_grads[n]
                    if gr is None or gr.abs().max().item() < 1e-5 * gmax:
                        continue
                    worst_l2 = max(worst_l2, ((tap[n].cpu() - gr).norm() / gr.norm()).item())
                assert worst_l2 < 3e-2, f"rank {rank} it {it}: averaged gradient L2 error {worst_l2}"   # LeakyReLU kink gate (tests/test_wrn_gpu.py)
                log.append(f"wrn rank {rank} it {it}: total {ld['train/total_loss']:.5f} (oracle {float(rec['total_loss']):.5f}) avg-grad L2 err {worst_l2:.2e} bn buffers {worst_b:.1e}")
            with torch.no_grad():
                for n, p in alg._net().named_parameters():
                    p.copy_(orc.p[n].detach())
                for n, p in alg.rewarder.named_parameters():
                    p.copy_(orc.rp[n].detach())
            alg._net().mark_weights_updated()
        flat = torch.cat([p.detach().flatten() for p in alg._net().parameters()] + [bf.detach().flatten().float() for bf in alg._net().buffers()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        q.put((rank, "ok", same, True, log))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:   # noqa: BLE001
        import traceback
        q.put((rank, "fail", traceback.format_exc(), None, []))
        raise


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_two_rank_wrn_syncbn_step_parity_vs_rank_oracles():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_wrn_worker, args=(r, world, 29657, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(120)
    for rank, status, same, _, log in sorted(res):
        assert status == "ok", same
        print("\n".join(log))
        assert same, "parameters / BatchNorm buffers differ between ranks after SyncBatchNorm + NCCL-averaged steps"
