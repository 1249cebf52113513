"""Plumbing shared by the native backbones (ViT, BERT): pooled workspaces and persistent buffers (stable device pointers are the
precondition of the engines' CUDA-graph replay), the flat gradient buffer whose pieces become `p.grad`, and the data-parallel
gradient exchange (one NCCL all-reduce(avg), or overlapped pieces as the backward's block ranges finish — DDP's buckets,
core/utils/misc.py:42-64)."""
from __future__ import annotations

import os

import torch

from .. import _lib as L


class NativeBackbone:
    """Mixin for an nn.Module; the module provides `_ordered_params()` (engine order) and, optionally, `_grad_params()` (flat
    gradient buffer order; default = engine order)."""

    def _init_native(self):
        self._keep_cache, self._ws_pool, self._bufs = {}, {}, {}
        # data parallel: block ranges of the backward whose gradients are all-reduced while the next range runs (<= 1: off); SRW_DP_SPLIT overrides
        self.dp_overlap_split = int(os.environ.get("SRW_DP_SPLIT", "4"))
        self._pending_reduce = []
        self._pa = self._pa_key = self._flat_grads = self._grad_views = self._ga = None

    # -- workspace pool / persistent buffers ------------------------------------------------------
    def _acquire_ws(self, wbytes, device):
        pool = self._ws_pool.setdefault((wbytes, str(device)), [])
        if pool:
            return pool.pop()
        # the engines want 1024-byte alignment (SWIZZLE_128B operand tiles inside the workspace); torch's caching allocator only
        # promises 512 for small blocks, so over-allocate and hand out an aligned view (the view keeps the storage alive)
        raw = torch.empty(wbytes + 1024, dtype=torch.uint8, device=device)
        off = (-raw.data_ptr()) % 1024
        return raw[off:off + wbytes]

    def _release_ws(self, ws):
        if ws is not None:
            self._ws_pool.setdefault((ws.numel(), str(ws.device)), []).append(ws)

    def release_pass(self, handle):
        self._release_ws(handle.pop("ws", None))

    def _buf(self, name, shape, device, dtype=torch.float32):
        key = (name, tuple(shape), str(device), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = self._bufs[key] = torch.empty(shape, dtype=dtype, device=device)
        return t

    def _native_params(self):
        ps = self._ordered_params()
        key = tuple(None if p is None else p.data_ptr() for p in ps)
        if self._pa_key != key:
            self._pa, self._pa_key = L.ptr_array(ps), key
        return ps, self._pa

    def _grad_params(self):
        return [p for p in self._ordered_params() if p is not None]

    def _ensure_flat_grads(self, dev):
        """One flat fp32 buffer for every gradient of the step (single all-reduce, single optimizer launch); `_grad_views[i]`
        is the view of `_grad_params()[i]`, `_ga` the engine-order pointer array (NULL for parameters without a gradient)."""
        if self._flat_grads is not None and self._flat_grads.device == dev:
            return
        gps = self._grad_params()
        self._flat_grads = torch.empty(sum(p.numel() for p in gps), dtype=torch.float32, device=dev)
        self._grad_views, off, by_id = [], 0, {}
        for p in gps:
            v = self._flat_grads[off:off + p.numel()].view_as(p)
            self._grad_views.append(v)
            by_id[id(p)] = v
            off += p.numel()
        self._ga = L.ptr_array([None if p is None else by_id.get(id(p)) for p in self._ordered_params()])

    # -- data parallel ----------------------------------------------------------------------------
    def _dp_bounds(self, depth):
        """Descending lower block bounds of all but the last range: dp_overlap_split = k -> k ranges of ~depth/k blocks."""
        k = int(self.dp_overlap_split)
        if k <= 1 or depth < 2:
            return []
        k = min(k, depth)
        return sorted({(depth * i) // k for i in range(1, k)} - {0}, reverse=True)

    @staticmethod
    def _allreduce_async(t, group):
        import torch.distributed as dist
        if dist.get_backend(group) == "nccl":
            return (dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group, async_op=True), None)
        return (dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True), (t, dist.get_world_size(group)))   # gloo: no AVG

    def allreduce_grads_(self):
        """Average the flat gradient over the data-parallel group (C1 in SURVEY.md §2.1); completes the overlapped
        all-reduces backward_native() already started, or runs one all-reduce of the whole buffer."""
        group = getattr(self, "_dp_group", None)
        if group is None or self._flat_grads is None:
            return
        pending, self._pending_reduce = getattr(self, "_pending_reduce", []), []
        if pending:
            for work, post in pending:
                work.wait()            # the current stream waits for NCCL's stream
                if post is not None:
                    post[0].div_(post[1])
            return
        from ..parallel import allreduce_mean_
        allreduce_mean_(self._flat_grads, group)
