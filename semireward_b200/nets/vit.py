"""B200-native VisionTransformer behind the reference's net-builder interface.

Mirrors semilearn/nets/vit/vit.py (interface only — all arithmetic runs in libsrw_b200.so):
  * builders  vit_tiny_patch2_32 / vit_small_patch2_32 / vit_small_patch16_224 / vit_base_patch16_96 /
    vit_base_patch16_224 (vit.py:323-408): f(pretrained=False, pretrained_path=None, **kw) -> nn.Module
  * module contract (vit.py:285-320): forward(x, only_fc=False, only_feat=False) -> {'logits','feat'}, extract(),
    no_weight_decay(), group_matcher(), num_features
  * identical state_dict keys / shapes / parameter registration order (152 tensors for ViT-S), so load_checkpoint,
    param_groups_layer_decay, EMA, DDP and torch.save keep working unchanged.
Parameters are ordinary fp32 nn.Parameters owned by PyTorch; the nn.Linear / nn.LayerNorm / nn.Conv2d children are
parameter holders only (same default initialisation as the reference under the same seed) and are never called.
There is no PyTorch fallback: without the CUDA library or on a CPU tensor forward() raises.
"""
from __future__ import annotations

import ctypes as C
from functools import partial

import torch
import torch.nn as nn

from .. import _lib as L
from ._native import NativeBackbone


class _Attention(nn.Module):  # holder for vit.py:78-107 parameters
    def __init__(self, dim, num_heads, qkv_bias=True):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):  # holder for vit.py:47-75 parameters
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):  # holder for vit.py:120-166 parameters
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, drop_path, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _Attention(dim, num_heads, qkv_bias)
        self.drop_path = float(drop_path)
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class _PatchEmbed(nn.Module):  # holder for vit.py:13-44 parameters
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class _VitFunction(torch.autograd.Function):
    """logits, feat = ViT(x) through srw_vit_forward / srw_vit_backward (include/srw.h)."""

    @staticmethod
    def forward(ctx, model, x, drop_scale, grad_batch, *params):
        lib = L.load()
        B = x.shape[0]
        cfg = model._cfg
        wbytes = lib.srw_vit_workspace_bytes(C.byref(cfg), B, grad_batch)
        if wbytes < 0:
            L.check(-2, "srw_vit_workspace_bytes")
        ws = model._acquire_ws(wbytes, x.device)
        logits = torch.empty(B, cfg.num_classes, dtype=torch.float32, device=x.device)
        feat = torch.empty(B, cfg.embed_dim, dtype=torch.float32, device=x.device)
        pa = L.ptr_array(params)
        a = L.VitFwdArgs(cfg=C.pointer(cfg), params=pa, weight_planes=model._weight_planes().data_ptr(), x=x.data_ptr(), batch=B,
                         grad_batch=grad_batch, drop_scale=L.ptr(drop_scale), logits=logits.data_ptr(), feat=feat.data_ptr(),
                         workspace=ws.data_ptr(), workspace_bytes=wbytes, gemm_impl=model.gemm_impl)
        L.check(lib.srw_vit_forward(C.byref(a), L.stream_ptr()), "srw_vit_forward")
        if grad_batch > 0:
            ctx.model, ctx.ws, ctx.wbytes, ctx.x, ctx.drop_scale = model, ws, wbytes, x, drop_scale
            ctx.B, ctx.grad_batch = B, grad_batch
            ctx.params = params
        else:
            model._release_ws(ws)   # stream-ordered: the next user of this workspace runs after this forward
        ctx.mark_non_differentiable()
        return logits, feat

    @staticmethod
    def backward(ctx, dlogits, dfeat):
        lib = L.load()
        model, params, Bg = ctx.model, ctx.params, ctx.grad_batch
        cfg = model._cfg
        dev = ctx.x.device
        numels = [p.numel() for p in params]
        flat = torch.empty(sum(numels), dtype=torch.float32, device=dev)
        grads, off = [], 0
        for p, n in zip(params, numels):
            grads.append(flat[off:off + n].view_as(p))
            off += n
        dl = dlogits[:Bg].contiguous() if dlogits is not None else torch.zeros(Bg, cfg.num_classes, device=dev)
        df = dfeat[:Bg].contiguous() if dfeat is not None else None
        a = L.VitBwdArgs(cfg=C.pointer(cfg), params=L.ptr_array(params), weight_planes=model._weight_planes().data_ptr(),
                         x=ctx.x.data_ptr(), batch=ctx.B, grad_batch=Bg, drop_scale=L.ptr(ctx.drop_scale), dlogits=dl.data_ptr(),
                         dfeat=L.ptr(df), grads=L.ptr_array(grads), accumulate_grads=0, workspace=ctx.ws.data_ptr(),
                         workspace_bytes=ctx.wbytes, gemm_impl=model.gemm_impl, block_lo=-1, block_hi=-1)
        L.check(lib.srw_vit_backward(C.byref(a), L.stream_ptr()), "srw_vit_backward")
        model._release_ws(ctx.ws)
        ctx.ws = None
        group = getattr(model, "_dp_group", None)
        if group is not None:   # data parallel: one all-reduce(avg) of the whole flat gradient (C1 in SURVEY.md §2.1)
            from ..parallel import allreduce_mean_
            allreduce_mean_(flat, group)
        return (None, None, None, None) + tuple(grads)


class VisionTransformer(NativeBackbone, nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, global_pool="token", embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0,
                 init_values=None, embed_layer=None, norm_layer=None, act_layer=None, block_fn=None):
        super().__init__()
        if global_pool != "token" or init_values is not None or drop_rate != 0.0 or attn_drop_rate != 0.0 or not qkv_bias:
            raise NotImplementedError("semireward_b200 ViT: only global_pool='token', no LayerScale, dropout 0, qkv_bias=True "
                                      "(the settings every SemiReward config uses) are built natively")
        if embed_layer is not None or norm_layer is not None or act_layer is not None or block_fn is not None:
            raise NotImplementedError("semireward_b200 ViT: custom embed/norm/act/block layers are not supported")
        norm_layer = partial(nn.LayerNorm, eps=1e-6)
        self.num_classes = num_classes
        self.global_pool = global_pool
        self.num_features = self.embed_dim = embed_dim
        self.num_tokens = 1
        self.patch_embed = _PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]  # vit.py:247-249
        self.blocks = nn.Sequential(*[_Block(embed_dim, num_heads, mlp_ratio, qkv_bias, dpr[i], norm_layer) for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)
        self.drop_path_rates = dpr
        self.gemm_impl = L.GEMM_TCGEN05
        self._cfg = L.VitConfig(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim, depth=depth,
                                num_heads=num_heads, hidden_dim=int(embed_dim * mlp_ratio), num_classes=num_classes, ln_eps=1e-6)
        self._planes = None
        self._planes_key = None
        self._init_native()

    # -- native plumbing ------------------------------------------------------------------------
    def _ordered_params(self):
        """Parameters in the order include/srw.h expects (== state_dict order)."""
        ps = [self.cls_token, self.pos_embed, self.patch_embed.proj.weight, self.patch_embed.proj.bias]
        for b in self.blocks:
            ps += [b.norm1.weight, b.norm1.bias, b.attn.qkv.weight, b.attn.qkv.bias, b.attn.proj.weight, b.attn.proj.bias,
                   b.norm2.weight, b.norm2.bias, b.mlp.fc1.weight, b.mlp.fc1.bias, b.mlp.fc2.weight, b.mlp.fc2.bias]
        ps += [self.norm.weight, self.norm.bias, self.head.weight, self.head.bias]
        return ps

    def _matrix_params(self):
        ps = [self.patch_embed.proj.weight]
        for b in self.blocks:
            ps += [b.attn.qkv.weight, b.attn.proj.weight, b.mlp.fc1.weight, b.mlp.fc2.weight]
        return ps

    def _weight_planes(self):
        """Split-bf16 cache of the GEMM weights, refreshed whenever a parameter was modified in place (optimizer step,
        load_state_dict, EMA copy) — detected through the tensors' version counters."""
        mats = self._matrix_params()
        key = tuple((p.data_ptr(), p._version) for p in mats)
        if self._planes is None or key != self._planes_key:
            lib = L.load()
            dev = mats[0].device
            if self._planes is None or self._planes.device != dev:
                nbytes = lib.srw_vit_weight_planes_bytes(C.byref(self._cfg))
                self._planes = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            params = [p.detach() for p in self._ordered_params()]
            L.check(lib.srw_vit_prepare_weights(C.byref(self._cfg), L.ptr_array(params), self._planes.data_ptr(), L.stream_ptr()),
                    "srw_vit_prepare_weights")
            self._planes_key = key
        return self._planes

    def mark_weights_updated(self, planes_fresh: bool = False):
        """Called by the fused optimizer: planes_fresh=True means it rewrote the weight planes itself."""
        if planes_fresh and self._planes is not None:
            self._planes_key = tuple((p.data_ptr(), p._version) for p in self._matrix_params())
        else:
            self._planes_key = None

    def _draw_drop_scale(self, batch, device, out=None):
        """[depth, 2, batch] DropPath multipliers mask/keep (timm semantics: per sample Bernoulli(keep)/keep)."""
        if not self.training or max(self.drop_path_rates) == 0.0:
            return None
        key = (batch, str(device))
        keep = self._keep_cache.get(key)
        if keep is None:
            keep = 1.0 - torch.tensor(self.drop_path_rates, dtype=torch.float32, device=device).view(-1, 1, 1)
            keep = self._keep_cache[key] = keep.expand(-1, 2, batch).contiguous()
        if out is None:
            out = torch.empty_like(keep)
        torch.bernoulli(keep, out=out)
        return out.div_(keep)

    # -- uniform surface the SSL step drives every native backbone through (nets/bert.py has the same four methods) ----
    def stochastic(self):
        return self.training and max(self.drop_path_rates) > 0.0

    def draw_streams(self, num_passes, nl, nu, device):
        """DropPath multipliers of `num_passes` backbone passes over (nl labelled, nu strong, nu weak) rows, drawn in the reference's
        order (one draw per pass over the whole concatenated batch) -> [depth, 2, pass, nl + 2 nu] in engine row order, or None."""
        if not self.stochastic():
            return None
        per = nl + 2 * nu
        return self._draw_drop_scale(num_passes * per, device).view(-1, 2, num_passes, per)

    def streams_for(self, draws, pieces, nl, nu, device):
        """pieces: [(pass, 'lb' | 's' | 'w'), ...] in launch row order -> [depth, 2, rows] multipliers for that launch."""
        if draws is None:
            return None
        sl = {"lb": slice(0, nl), "s": slice(nl, nl + nu), "w": slice(nl + nu, nl + 2 * nu)}
        cols = [draws[:, :, ps_, sl[part]] for ps_, part in pieces]
        return (cols[0] if len(cols) == 1 else torch.cat(cols, dim=2)).contiguous()

    def concat_inputs(self, parts, device):
        rows = sum(p.shape[0] for p in parts)
        xb = self.input_buffer((rows,) + tuple(parts[0].shape[1:]), device)
        torch.cat(list(parts), out=xb)
        return xb

    # -- persistent I/O buffers: stable device pointers let the engine replay CUDA graphs (include/srw.h) ----
    def input_buffer(self, shape, device):
        """Persistent staging buffer for the batch: `torch.cat(parts, out=net.input_buffer(...))` then forward_native()."""
        return self._buf("x", shape, device)

    @torch.no_grad()
    def forward_native(self, x, grad_batch=0, drop_scale=None):
        """One autograd-free forward through srw_vit_forward on persistent buffers -> (logits, feat, handle).  `handle`
        keeps the activations of the first `grad_batch` rows for backward_native(); release it with release_pass() when
        no backward will follow.  logits / feat are fresh tensors."""
        if not x.is_cuda:
            raise RuntimeError("semireward_b200 ViT runs on CUDA (sm_100a) only; there is no CPU path")
        lib, cfg, dev, B = L.load(), self._cfg, x.device, x.shape[0]
        xb = self.input_buffer(x.shape, dev)
        if x.data_ptr() != xb.data_ptr():
            xb.copy_(x)
        params, pa = self._native_params()
        wbytes = lib.srw_vit_workspace_bytes(C.byref(cfg), B, grad_batch)
        if wbytes < 0:
            L.check(-2, "srw_vit_workspace_bytes")
        ws = self._acquire_ws(wbytes, dev)
        ds = drop_scale
        if self.training and max(self.drop_path_rates) > 0.0:
            # one persistent buffer per workspace: a pass kept alive for backward keeps its own DropPath draw; explicit
            # multipliers are copied into it so the engine sees a stable pointer (graph replay)
            buf = self._buf(("drop", ws.data_ptr()), (cfg.depth, 2, B), dev)
            if ds is None:
                ds = self._draw_drop_scale(B, dev, out=buf)
            elif ds.data_ptr() != buf.data_ptr():
                buf.copy_(ds)
                ds = buf
        lo, fe = self._buf("logits", (B, cfg.num_classes), dev), self._buf("feat", (B, cfg.embed_dim), dev)
        a = L.VitFwdArgs(cfg=C.pointer(cfg), params=pa, weight_planes=self._weight_planes().data_ptr(), x=xb.data_ptr(), batch=B,
                         grad_batch=grad_batch, drop_scale=L.ptr(ds), logits=lo.data_ptr(), feat=fe.data_ptr(),
                         workspace=ws.data_ptr(), workspace_bytes=wbytes, gemm_impl=self.gemm_impl)
        L.check(lib.srw_vit_forward(C.byref(a), L.stream_ptr()), "srw_vit_forward")
        handle = dict(ws=ws, wbytes=wbytes, x=xb, drop_scale=ds, B=B, grad_batch=grad_batch)
        if grad_batch == 0:
            self.release_pass(handle)
        return lo.clone(), fe.clone(), handle

    def dlogits_buffer(self, grad_batch, device):
        return self._buf("dlogits", (grad_batch, self._cfg.num_classes), device)

    @torch.no_grad()
    def backward_native(self, handle, dlogits, dfeat=None, accumulate=False, final=True):
        """srw_vit_backward for a pass of forward_native(): parameter gradients (state_dict order) into the persistent flat
        buffer -> (flat, views).  Data parallel: the flat buffer is all-reduced (mean) here unless more passes accumulate."""
        lib, cfg = L.load(), self._cfg
        Bg, dev = handle["grad_batch"], dlogits.device
        params, pa = self._native_params()
        self._ensure_flat_grads(dev)
        dl = self.dlogits_buffer(Bg, dev)
        if dlogits.data_ptr() != dl.data_ptr():
            dl.copy_(dlogits)
        df = None
        if dfeat is not None:
            df = self._buf("dfeat", (Bg, cfg.embed_dim), dev)
            df.copy_(dfeat)
        a = L.VitBwdArgs(cfg=C.pointer(cfg), params=pa, weight_planes=self._weight_planes().data_ptr(), x=handle["x"].data_ptr(),
                         batch=handle["B"], grad_batch=Bg, drop_scale=L.ptr(handle["drop_scale"]), dlogits=dl.data_ptr(), dfeat=L.ptr(df),
                         grads=self._ga, accumulate_grads=int(bool(accumulate)), workspace=handle["ws"].data_ptr(),
                         workspace_bytes=handle["wbytes"], gemm_impl=self.gemm_impl, block_lo=-1, block_hi=-1)
        group = getattr(self, "_dp_group", None)
        self._pending_reduce = []
        bounds = self._dp_bounds(cfg.depth) if (group is not None and final) else []
        if bounds:
            # data parallel with overlap (DDP's bucketed all-reduce, core/utils/misc.py:42-64): the backward runs as descending
            # block ranges; the gradients of a finished range (and, for the first one, of the final norm and the head) are a
            # contiguous piece of the flat buffer, all-reduced on NCCL's stream while the next range is computed
            hi, end = cfg.depth - 1, self._flat_grads.numel()
            for lo in bounds + [0]:
                a.block_lo, a.block_hi = lo, hi
                L.check(lib.srw_vit_backward(C.byref(a), L.stream_ptr()), "srw_vit_backward")
                off = sum(p.numel() for p in params[:4 + 12 * lo]) if lo > 0 else 0
                self._pending_reduce.append(self._allreduce_async(self._flat_grads[off:end], group))
                hi, end = lo - 1, off
        else:
            L.check(lib.srw_vit_backward(C.byref(a), L.stream_ptr()), "srw_vit_backward")
        self.release_pass(handle)
        return self._flat_grads, self._grad_views

    def _run(self, x, grad_batch=None, drop_scale=None):
        if not x.is_cuda:
            raise RuntimeError("semireward_b200 ViT runs on CUDA (sm_100a) only; there is no CPU path")
        params = self._ordered_params()
        x = x.contiguous().float()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        gb = (x.shape[0] if grad_batch is None else int(grad_batch)) if need_grad else 0
        if drop_scale is None:
            drop_scale = self._draw_drop_scale(x.shape[0], x.device)
        return _VitFunction.apply(self, x, drop_scale, gb, *params)

    # -- reference interface --------------------------------------------------------------------
    @torch.no_grad()
    def extract(self, x):
        """vit.py:277-283: every token after the final LayerNorm, [B, N, D].  Inference output of the fused engine (the engine's
        backward starts from logits / feat, so no gradient flows through this tensor)."""
        if not x.is_cuda:
            raise RuntimeError("semireward_b200 ViT runs on CUDA (sm_100a) only; there is no CPU path")
        lib, cfg, dev, B = L.load(), self._cfg, x.device, x.shape[0]
        x = x.contiguous().float()
        params, pa = self._native_params()
        wbytes = lib.srw_vit_workspace_bytes(C.byref(cfg), B, 0)
        ws = self._acquire_ws(wbytes, dev)
        ntok = self.patch_embed.num_patches + 1
        tokens = torch.empty(B, ntok, cfg.embed_dim, dtype=torch.float32, device=dev)
        lo, fe = torch.empty(B, cfg.num_classes, device=dev), torch.empty(B, cfg.embed_dim, device=dev)
        ds = self._draw_drop_scale(B, dev)
        a = L.VitFwdArgs(cfg=C.pointer(cfg), params=pa, weight_planes=self._weight_planes().data_ptr(), x=x.data_ptr(), batch=B,
                         grad_batch=0, drop_scale=L.ptr(ds), logits=lo.data_ptr(), feat=fe.data_ptr(), workspace=ws.data_ptr(),
                         workspace_bytes=wbytes, gemm_impl=self.gemm_impl, tokens_out=tokens.data_ptr())
        L.check(lib.srw_vit_forward(C.byref(a), L.stream_ptr()), "srw_vit_forward")
        self._release_ws(ws)
        return tokens

    def forward(self, x, only_fc=False, only_feat=False, grad_batch=None, drop_scale=None, **kwargs):
        """grad_batch (extension): only the first `grad_batch` rows are back-propagated (rows after it must only feed
        detached consumers, like the weak-augmentation rows of the SSL batch)."""
        if only_fc:
            # vit.py:293-294 / eval.py:82-83: the classifier alone on externally pooled features [B, D].  Off the train-step hot
            # path (a [B, D] x [D, C] product), so it is the plain library linear on the module's own parameters.
            return torch.nn.functional.linear(x, self.head.weight, self.head.bias)
        logits, feat = self._run(x, grad_batch, drop_scale)
        if only_feat:
            return feat
        return {"logits": logits, "feat": feat}

    def no_weight_decay(self):
        return {"pos_embed", "cls_token"}

    def group_matcher(self, coarse=False, prefix=""):
        return dict(stem=r"^{}cls_token|{}pos_embed|{}patch_embed".format(prefix, prefix, prefix),
                    blocks=[(r"^{}blocks\.(\d+)".format(prefix), None), (r"^{}norm".format(prefix), (99999,))])


def _build(defaults, pretrained, pretrained_path, kwargs):
    kw = dict(defaults)
    kw.update(kwargs)
    model = VisionTransformer(**kw)
    if pretrained:   # vit.py:352-354 and twins
        from .utils import load_checkpoint
        model = load_checkpoint(model, pretrained_path)
    return model


def vit_tiny_patch2_32(pretrained=False, pretrained_path=None, **kwargs):
    return _build(dict(img_size=32, patch_size=2, embed_dim=192, depth=12, num_heads=3, drop_path_rate=0.1), pretrained, pretrained_path, kwargs)


def vit_small_patch2_32(pretrained=False, pretrained_path=None, **kwargs):
    return _build(dict(img_size=32, patch_size=2, embed_dim=384, depth=12, num_heads=6, drop_path_rate=0.2), pretrained, pretrained_path, kwargs)


def vit_small_patch16_224(pretrained=False, pretrained_path=None, **kwargs):
    return _build(dict(patch_size=16, embed_dim=384, depth=12, num_heads=6, drop_path_rate=0.2), pretrained, pretrained_path, kwargs)


def vit_base_patch16_96(pretrained=False, pretrained_path=None, **kwargs):
    return _build(dict(img_size=96, patch_size=16, embed_dim=768, depth=12, num_heads=12, drop_path_rate=0.2), pretrained, pretrained_path, kwargs)


def vit_base_patch16_224(pretrained=False, pretrained_path=None, **kwargs):
    return _build(dict(patch_size=16, embed_dim=768, depth=12, num_heads=12, drop_path_rate=0.2), pretrained, pretrained_path, kwargs)
