"""B200-native ClassificationHubert behind the reference's net-builder interface (semilearn/nets/hubert/hubert.py:10-63).

  * builder `hubert_base` (hubert.py:59-61): f(pretrained=False, pretrained_path=None, **kw) -> nn.Module
  * module contract (hubert.py:24-57): forward(x = waveform [B, T] fp32, only_fc=False, only_feat=False) -> {'logits', 'feat'},
    extract(), group_matcher(), no_weight_decay(), num_features = 768
  * identical state_dict keys / shapes / registration order as the reference's module (Hugging Face `HubertModel` under `model.`,
    the weight-normalised positional conv as `parametrizations.weight.original0 / original1`, then `classifier.0`, `classifier.2`):
    211 tensors for hubert-base.
The reference builds the encoder with `HubertModel.from_pretrained('facebook/hubert-base-ls960')` (a hub download); offline, and for
random-init synthetic runs, the builder initialises like `HubertModel(HubertConfig())` and loads
`pretrained_path` when it names a local state-dict file.  All arithmetic runs in libsrw_b200.so (srw_hubert_forward /
srw_hubert_backward); the nn.Module children are parameter holders and are never called.  There is no PyTorch fallback.

Randomness of a train-mode call.  The reference draws, per model call: nn.Dropout masks (feature projection, encoder input, attention
probabilities, attention / FFN outputs, FFN activation, the wrapper's own dropout), one LayerDrop coin per encoder layer, and the
SpecAugment spans (`_compute_mask_indices`, numpy's global RNG).  Here dropout is counter-based (include/srw.h: srw_dropout; stream key
per call = call_key(dropout_seed, call index)); LayerDrop coins and SpecAugment spans are drawn on the host by `draw_streams` from the
module's own numpy Generator (`stochastic_seed`) with the reference's distributions, and handed to the engine as explicit inputs."""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L
from ._native import NativeBackbone
from .bert import call_key

CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)


class _ConvLayer(nn.Module):
    def __init__(self, cin, cout, k, group_norm):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, k, bias=False)
        if group_norm:
            self.layer_norm = nn.GroupNorm(cout, cout, affine=True)


class _FeatureExtractor(nn.Module):
    def __init__(self, dim, kernels):
        super().__init__()
        self.conv_layers = nn.ModuleList([_ConvLayer(1 if i == 0 else dim, dim, k, i == 0) for i, k in enumerate(kernels)])


class _FeatureProjection(nn.Module):
    def __init__(self, dim, H, eps):
        super().__init__()
        self.layer_norm = nn.LayerNorm(dim, eps=eps)
        self.projection = nn.Linear(dim, H)


class _WeightNormParams(nn.Module):   # torch.nn.utils.parametrizations.weight_norm(conv, dim=2): original0 = g [1, 1, K], original1 = v
    def __init__(self, H, gc, K):
        super().__init__()
        self.original0 = nn.Parameter(torch.ones(1, 1, K))
        self.original1 = nn.Parameter(torch.zeros(H, gc, K))


class _Parametrizations(nn.Module):
    def __init__(self, H, gc, K):
        super().__init__()
        self.weight = _WeightNormParams(H, gc, K)


class _PosConv(nn.Module):
    def __init__(self, H, gc, K):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(H))
        self.parametrizations = _Parametrizations(H, gc, K)


class _PosConvEmbed(nn.Module):
    def __init__(self, H, groups, K):
        super().__init__()
        self.conv = _PosConv(H, H // groups, K)


class _HubAttention(nn.Module):
    def __init__(self, H):
        super().__init__()
        self.k_proj, self.v_proj, self.q_proj, self.out_proj = nn.Linear(H, H), nn.Linear(H, H), nn.Linear(H, H), nn.Linear(H, H)


class _FeedForward(nn.Module):
    def __init__(self, H, I):
        super().__init__()
        self.intermediate_dense = nn.Linear(H, I)
        self.output_dense = nn.Linear(I, H)


class _HubLayer(nn.Module):
    def __init__(self, H, I, eps):
        super().__init__()
        self.attention = _HubAttention(H)
        self.layer_norm = nn.LayerNorm(H, eps=eps)
        self.feed_forward = _FeedForward(H, I)
        self.final_layer_norm = nn.LayerNorm(H, eps=eps)


class _HubEncoder(nn.Module):
    def __init__(self, H, I, n, eps, groups, K):
        super().__init__()
        self.pos_conv_embed = _PosConvEmbed(H, groups, K)
        self.layer_norm = nn.LayerNorm(H, eps=eps)
        self.layers = nn.ModuleList([_HubLayer(H, I, eps) for _ in range(n)])


class _HubertModel(nn.Module):   # parameter holder with Hugging Face HubertModel's module tree
    def __init__(self, H, I, n, eps, dim, kernels, groups, K):
        super().__init__()
        self.masked_spec_embed = nn.Parameter(torch.empty(H).uniform_())
        self.feature_extractor = _FeatureExtractor(dim, kernels)
        self.feature_projection = _FeatureProjection(dim, H, eps)
        self.encoder = _HubEncoder(H, I, n, eps, groups, K)


class ClassificationHubert(NativeBackbone, nn.Module):
    def __init__(self, name="facebook/hubert-base-ls960", num_classes=2, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, conv_dim=512, conv_kernel=CONV_KERNEL, conv_stride=CONV_STRIDE, num_conv_pos_embeddings=128,
                 num_conv_pos_embedding_groups=16, layer_norm_eps=1e-5, feat_proj_dropout=0.1, hidden_dropout=0.1, attention_dropout=0.1,
                 activation_dropout=0.1, pooled_dropout=0.1, layerdrop=0.1, apply_spec_augment=True, mask_time_prob=0.05, mask_time_length=10,
                 mask_time_min_masks=2, initializer_range=0.02):
        super().__init__()
        H = hidden_size
        self.model = _HubertModel(H, intermediate_size, num_hidden_layers, layer_norm_eps, conv_dim, conv_kernel, num_conv_pos_embedding_groups,
                                  num_conv_pos_embeddings)
        self.dropout = nn.Dropout(p=pooled_dropout, inplace=False)   # holder of p (hubert.py:15); the mask is drawn inside the engine
        self.num_features = H
        self.classifier = nn.Sequential(nn.Linear(H, H), nn.GELU(), nn.Linear(H, num_classes))
        with torch.no_grad():   # HubertPreTrainedModel._init_weights; the classifier keeps nn.Linear's default like the reference
            for m in self.model.modules():
                if isinstance(m, nn.Linear):
                    m.weight.normal_(0.0, initializer_range)
                    m.bias.zero_()
                elif isinstance(m, nn.Conv1d):
                    nn.init.kaiming_normal_(m.weight)
            fp = self.model.feature_projection.projection
            k = math.sqrt(1.0 / fp.in_features)
            fp.weight.uniform_(-k, k)
            fp.bias.uniform_(-k, k)
            pc = self.model.encoder.pos_conv_embed.conv
            v = pc.parametrizations.weight.original1
            v.normal_(0.0, 2.0 * math.sqrt(1.0 / (num_conv_pos_embeddings * H)))
            pc.parametrizations.weight.original0.copy_(v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt())   # weight_norm init: g = ||v||
        self.gemm_impl = L.GEMM_TCGEN05
        self._cfg = L.HubertConfig(hidden=H, layers=num_hidden_layers, heads=num_attention_heads, intermediate=intermediate_size, num_classes=num_classes,
                                   conv_dim=conv_dim, num_conv=len(conv_kernel), pos_kernel=num_conv_pos_embeddings, pos_groups=num_conv_pos_embedding_groups,
                                   ln_eps=layer_norm_eps, p_feat_proj=feat_proj_dropout, p_hidden=hidden_dropout, p_attn=attention_dropout,
                                   p_act=activation_dropout, p_pooled=pooled_dropout)
        for i, (k_, s_) in enumerate(zip(conv_kernel, conv_stride)):
            self._cfg.conv_kernel[i], self._cfg.conv_stride[i] = k_, s_
        self.layerdrop, self.apply_spec_augment = layerdrop, apply_spec_augment
        self.mask_time_prob, self.mask_time_length, self.mask_time_min_masks = mask_time_prob, mask_time_length, mask_time_min_masks
        self._planes = self._planes_key = None
        self._front_stale = False
        self._init_native()
        self.dropout_seed = 0          # seed of the counter-based dropout streams (call_key)
        self.stochastic_seed = 0       # seed of the host-side LayerDrop / SpecAugment draws
        self._rng = None
        self._calls = 0                # model calls made so far: every call of the reference draws fresh randomness
        self._call_draws = {}          # call index -> (layer_skip uint8 [layers], mask_time bool [n, F] or None)
        self.depth = num_hidden_layers

    # -- native plumbing ------------------------------------------------------------------------
    def _ordered_params(self):
        """Engine order == state_dict order (include/srw.h)."""
        m = self.model
        ps = [m.masked_spec_embed]
        for i, cl in enumerate(m.feature_extractor.conv_layers):
            ps.append(cl.conv.weight)
            if i == 0:
                ps += [cl.layer_norm.weight, cl.layer_norm.bias]
        fp, enc = m.feature_projection, m.encoder
        pc = enc.pos_conv_embed.conv
        ps += [fp.layer_norm.weight, fp.layer_norm.bias, fp.projection.weight, fp.projection.bias, pc.bias, pc.parametrizations.weight.original0,
               pc.parametrizations.weight.original1, enc.layer_norm.weight, enc.layer_norm.bias]
        for ly in enc.layers:
            a, ff = ly.attention, ly.feed_forward
            ps += [a.k_proj.weight, a.k_proj.bias, a.v_proj.weight, a.v_proj.bias, a.q_proj.weight, a.q_proj.bias, a.out_proj.weight, a.out_proj.bias,
                   ly.layer_norm.weight, ly.layer_norm.bias, ff.intermediate_dense.weight, ff.intermediate_dense.bias, ff.output_dense.weight,
                   ff.output_dense.bias, ly.final_layer_norm.weight, ly.final_layer_norm.bias]
        ps += [self.classifier[0].weight, self.classifier[0].bias, self.classifier[2].weight, self.classifier[2].bias]
        return ps

    def _grad_params(self):
        """Flat gradient buffer order: d(q | k | v).weight contiguous in that order and likewise their biases (one packed projection
        GEMM, include/srw.h)."""
        ps = []
        m = self.model
        ps.append(m.masked_spec_embed)
        for i, cl in enumerate(m.feature_extractor.conv_layers):
            ps.append(cl.conv.weight)
            if i == 0:
                ps += [cl.layer_norm.weight, cl.layer_norm.bias]
        fp, enc = m.feature_projection, m.encoder
        pc = enc.pos_conv_embed.conv
        ps += [fp.layer_norm.weight, fp.layer_norm.bias, fp.projection.weight, fp.projection.bias, pc.bias, pc.parametrizations.weight.original0,
               pc.parametrizations.weight.original1, enc.layer_norm.weight, enc.layer_norm.bias]
        for ly in enc.layers:
            a, ff = ly.attention, ly.feed_forward
            ps += [a.q_proj.weight, a.k_proj.weight, a.v_proj.weight, a.q_proj.bias, a.k_proj.bias, a.v_proj.bias, a.out_proj.weight, a.out_proj.bias,
                   ly.layer_norm.weight, ly.layer_norm.bias, ff.intermediate_dense.weight, ff.intermediate_dense.bias, ff.output_dense.weight,
                   ff.output_dense.bias, ly.final_layer_norm.weight, ly.final_layer_norm.bias]
        ps += [self.classifier[0].weight, self.classifier[0].bias, self.classifier[2].weight, self.classifier[2].bias]
        return ps

    def _layer_offset(self, lo):
        gps = self._grad_params()
        n_front = len(gps) - 16 * self.depth - 4
        return sum(p.numel() for p in gps[:n_front + 16 * lo])

    def _weight_planes(self):
        ps = self._ordered_params()
        key = tuple((p.data_ptr(), p._version) for p in ps)
        lib = L.load()
        if self._planes is None or key != self._planes_key:
            dev = ps[0].device
            if self._planes is None or self._planes.device != dev:
                self._planes = torch.empty(lib.srw_hubert_weight_planes_bytes(C.byref(self._cfg)), dtype=torch.uint8, device=dev)
            L.check(lib.srw_hubert_prepare_weights(C.byref(self._cfg), L.ptr_array([p.detach() for p in ps]), self._planes.data_ptr(), L.stream_ptr()),
                    "srw_hubert_prepare_weights")
            self._planes_key, self._front_stale = key, False
        elif self._front_stale:   # the fused optimizer refreshed the encoder matrices' planes itself; the re-laid-out operands follow here
            L.check(lib.srw_hubert_prepare_front(C.byref(self._cfg), L.ptr_array([p.detach() for p in ps]), self._planes.data_ptr(), L.stream_ptr()),
                    "srw_hubert_prepare_front")
            self._front_stale = False
        return self._planes

    def weight_plane_slot(self, idx):
        """(byte offset, cols, ldp, plane stride) of parameter `idx` (engine order) inside the plane cache for the encoder matrices, None
        for everything whose operand is a re-layout (conv stem, weight-normalised positional taps) or has no planes."""
        off, cols, ldp, ps_ = L.i64(), L.i32(), L.i32(), L.i64()
        if L.load().srw_hubert_weight_plane_slot(C.byref(self._cfg), idx, C.byref(off), C.byref(cols), C.byref(ldp), C.byref(ps_)) == 0:
            return off.value, cols.value, ldp.value, ps_.value
        return None

    def mark_weights_updated(self, planes_fresh: bool = False):
        if planes_fresh and self._planes is not None:   # FusedAdamW: parameters and the slotted planes were written by the same kernel
            self._planes_key = tuple((p.data_ptr(), p._version) for p in self._ordered_params())
            self._front_stale = True
        else:
            self._planes_key = None

    def stochastic(self):
        c = self._cfg
        return self.training and (max(c.p_feat_proj, c.p_hidden, c.p_attn, c.p_act, c.p_pooled) > 0.0 or self.layerdrop > 0.0 or
                                  (self.apply_spec_augment and self.mask_time_prob > 0.0))

    def frames(self, samples):
        return L.load().srw_hubert_frames(C.byref(self._cfg), int(samples))

    # -- per-call randomness -----------------------------------------------------------------------------------------------------
    def _spec_mask(self, n, F, rng):
        """SpecAugment spans with the distribution of transformers' _compute_mask_indices (no attention mask: every clip has F
        frames): per clip, num_spans = int(mask_prob * F / span + U[0, 1)) clamped to >= min_masks and to what fits; span starts
        drawn without replacement from [0, F - span]; overlapping spans merge."""
        span, prob = self.mask_time_length, self.mask_time_prob
        if not self.apply_spec_augment or prob <= 0.0 or span > F:
            return None
        m = np.zeros((n, F), dtype=bool)
        for b in range(n):
            k = int(prob * F / span + rng.random())
            k = max(k, self.mask_time_min_masks)
            if k * span > F:
                k = F // span
            if F - (span - 1) < k:
                k = max(F - (span - 1), 0)
            starts = rng.choice(F - (span - 1), k, replace=False) if k > 0 else np.zeros(0, dtype=np.int64)
            for s0 in starts:
                m[b, s0:s0 + span] = True
        return m

    def draw_streams(self, num_passes, nl, nu, device):
        """Reserve the randomness of `num_passes` passes of three model calls each (labelled, strong, weak: the reference's call
        order, srflexmatch.py:119-130) -> first call index.  The LayerDrop coins and SpecAugment masks of every call are drawn here,
        in call order."""
        first = self._calls
        self._calls += 3 * num_passes
        if self.training and (self.layerdrop > 0.0 or (self.apply_spec_augment and self.mask_time_prob > 0.0)):
            if self._rng is None:
                self._rng = np.random.default_rng(self.stochastic_seed)
            spec = self.apply_spec_augment and self.mask_time_prob > 0.0
            for c in range(first, first + 3 * num_passes):
                skip = (self._rng.random(self.depth) < self.layerdrop).astype(np.uint8) if self.layerdrop > 0.0 else np.zeros(self.depth, dtype=np.uint8)
                # the frame count is only known when the waveforms arrive: the call keeps its own seed and forward_native draws the spans
                self._call_draws[c] = (skip, int(self._rng.integers(0, 2 ** 63 - 1)) if spec else None)
        return first

    def set_call_draws(self, call, layer_skip=None, mask_time=None):
        """Inject the LayerDrop / SpecAugment decisions of one model call (tests: the same decisions go into the oracle)."""
        skip = np.zeros(self.depth, dtype=np.uint8) if layer_skip is None else np.asarray(layer_skip, dtype=np.uint8)
        self._call_draws[call] = (skip, None if mask_time is None else np.asarray(mask_time, dtype=bool))

    def streams_for(self, draws, pieces, nl, nu, device):
        """pieces: [(pass, 'lb' | 's' | 'w'), ...] in launch row order -> dict(keys, rows, segments, skip, mask) or None."""
        if not self.stochastic():
            return None
        keys, rows, seg, skips, masks = [], [], [0], [], []
        for ps_, part in pieces:
            n = nl if part == "lb" else nu
            call = draws + 3 * ps_ + {"lb": 0, "s": 1, "w": 2}[part]
            keys += [call_key(self.dropout_seed, call)] * n
            rows += list(range(n))
            seg.append(seg[-1] + n)
            skip, mask = self._call_draws.get(call, (np.zeros(self.depth, dtype=np.uint8), None))
            skips.append(skip)
            masks.append(mask)     # None, an explicit bool [n, F] array, or the seed the spans are drawn from
        for ps_, part in pieces:   # a call's draws are used by exactly one launch
            self._call_draws.pop(draws + 3 * ps_ + {"lb": 0, "s": 1, "w": 2}[part], None)
        c = self._cfg
        has_drop = max(c.p_feat_proj, c.p_hidden, c.p_attn, c.p_act, c.p_pooled) > 0.0
        out = dict(keys=None, rows=None, segments=np.asarray(seg, dtype=np.int32), skip=np.stack(skips).astype(np.uint8), masks=masks)
        if has_drop:
            out["keys"] = torch.from_numpy(np.asarray(keys, dtype=np.uint32).view(np.int32).copy())
            out["rows"] = torch.tensor(rows, dtype=torch.int32)
        return out

    def _mask_time(self, spec, F):
        """uint8 [S, F] SpecAugment mask of a launch from the per-call entries of streams_for(), or None."""
        seg, masks = spec["segments"], spec["masks"]
        if all(m is None for m in masks):
            return None
        full = np.zeros((int(seg[-1]), F), dtype=np.uint8)
        for i, m in enumerate(masks):
            n = int(seg[i + 1] - seg[i])
            if m is None or n == 0:
                continue
            if not isinstance(m, np.ndarray):
                m = self._spec_mask(n, F, np.random.default_rng(m))
            if m is not None:
                full[seg[i]:seg[i + 1]] = m
        return torch.from_numpy(full)

    def concat_inputs(self, parts, device):
        """parts: waveforms [n_i, T] -> one persistent [S, T] buffer on `device`.  The three calls of a step must have the same
        clip length (the reference's loaders pad every batch to the configured max_length_seconds)."""
        Ts = {int(p.shape[1]) for p in parts}
        if len(Ts) != 1:
            raise ValueError(f"native HuBERT: the calls of one launch must share the clip length (got {sorted(Ts)} samples)")
        T = Ts.pop()
        S = sum(int(p.shape[0]) for p in parts)
        buf = self._buf("wav", (S, T), device)
        torch.cat([p.to(torch.float32) for p in parts], out=buf)
        return buf

    # -- engine calls ---------------------------------------------------------------------------
    @torch.no_grad()
    def forward_native(self, x, grad_batch=0, drop_scale=None):
        """One autograd-free forward through srw_hubert_forward -> (logits, feat, handle).  x = waveform buffer from concat_inputs();
        drop_scale = streams_for(...) or None (no dropout / LayerDrop / SpecAugment)."""
        wav = x
        if not wav.is_cuda:
            raise RuntimeError("semireward_b200 HuBERT runs on CUDA (sm_100a) only; there is no CPU path")
        lib, cfg, dev = L.load(), self._cfg, wav.device
        S, T = wav.shape
        params, pa = self._native_params()
        wbytes = lib.srw_hubert_workspace_bytes(C.byref(cfg), S, T, grad_batch)
        if wbytes < 0:
            L.check(-2, "srw_hubert_workspace_bytes")
        ws = self._acquire_ws(wbytes, dev)
        kt = rt = mk = None
        seg = skip = None
        nseg = 0
        if drop_scale is not None:
            if drop_scale["keys"] is not None:
                kt, rt = self._buf(("dkey", ws.data_ptr()), (S,), dev, torch.int32), self._buf(("drow", ws.data_ptr()), (S,), dev, torch.int32)
                kt.copy_(drop_scale["keys"], non_blocking=True)
                rt.copy_(drop_scale["rows"], non_blocking=True)
            mt = self._mask_time(drop_scale, self.frames(T))
            if mt is not None:
                mk = self._buf(("mask_time", ws.data_ptr()), tuple(mt.shape), dev, torch.uint8)
                mk.copy_(mt.pin_memory(), non_blocking=True)
            if drop_scale["skip"].any():
                seg, skip = np.ascontiguousarray(drop_scale["segments"]), np.ascontiguousarray(drop_scale["skip"])
                nseg = len(seg) - 1
        lo, fe = self._buf("logits", (S, cfg.num_classes), dev), self._buf("feat", (S, cfg.hidden), dev)
        a = L.HubertFwdArgs(cfg=C.pointer(cfg), params=pa, weight_planes=self._weight_planes().data_ptr(), wav=wav.data_ptr(), ld_wav=wav.stride(0),
                            batch=S, samples=T, grad_batch=grad_batch, mask_time=L.ptr(mk), drop_seq_key=L.ptr(kt), drop_seq_row=L.ptr(rt),
                            num_segments=nseg, segment_start=None if seg is None else seg.ctypes.data, layer_skip=None if skip is None else skip.ctypes.data,
                            logits=lo.data_ptr(), feat=fe.data_ptr(), workspace=ws.data_ptr(), workspace_bytes=wbytes, gemm_impl=self.gemm_impl)
        L.check(lib.srw_hubert_forward(C.byref(a), L.stream_ptr()), "srw_hubert_forward")
        handle = dict(ws=ws, wbytes=wbytes, x=wav, keys=kt, rows=rt, mask=mk, seg=seg, skip=skip, nseg=nseg, B=S, T=T, grad_batch=grad_batch)
        if grad_batch == 0:
            self.release_pass(handle)
        return lo.clone(), fe.clone(), handle

    def dlogits_buffer(self, grad_batch, device):
        return self._buf("dlogits", (grad_batch, self._cfg.num_classes), device)

    @torch.no_grad()
    def backward_native(self, handle, dlogits, dfeat=None, accumulate=False, final=True):
        lib, cfg = L.load(), self._cfg
        Sg, dev = handle["grad_batch"], dlogits.device
        params, pa = self._native_params()
        self._ensure_flat_grads(dev)
        dl = self.dlogits_buffer(Sg, dev)
        if dlogits.data_ptr() != dl.data_ptr():
            dl.copy_(dlogits)
        df = None
        if dfeat is not None:
            df = self._buf("dfeat", (Sg, cfg.hidden), dev)
            df.copy_(dfeat)
        wav, seg, skip = handle["x"], handle["seg"], handle["skip"]
        a = L.HubertBwdArgs(cfg=C.pointer(cfg), params=pa, weight_planes=self._weight_planes().data_ptr(), wav=wav.data_ptr(), ld_wav=wav.stride(0),
                            batch=handle["B"], samples=handle["T"], grad_batch=Sg, mask_time=L.ptr(handle["mask"]), drop_seq_key=L.ptr(handle["keys"]),
                            drop_seq_row=L.ptr(handle["rows"]), num_segments=handle["nseg"], segment_start=None if seg is None else seg.ctypes.data,
                            layer_skip=None if skip is None else skip.ctypes.data, dlogits=dl.data_ptr(), dfeat=L.ptr(df), grads=self._ga,
                            accumulate_grads=int(bool(accumulate)), workspace=handle["ws"].data_ptr(), workspace_bytes=handle["wbytes"],
                            gemm_impl=self.gemm_impl)
        L.check(lib.srw_hubert_backward(C.byref(a), L.stream_ptr()), "srw_hubert_backward")
        self._pending_reduce = []
        self.release_pass(handle)
        return self._flat_grads, self._grad_views

    # -- reference interface --------------------------------------------------------------------
    @torch.no_grad()
    def _infer(self, x):
        dev = x.device
        n = x.shape[0]
        inp = self.concat_inputs([x], dev)
        spec = None
        if self.stochastic():   # a train-mode call outside the SSL step still draws its randomness, like the reference's module would
            d = self.draw_streams(1, n, 0, dev)
            spec = self.streams_for(d, [(0, "lb")], n, 0, dev)
        lg, ft, _ = self.forward_native(inp, grad_batch=0, drop_scale=spec)
        return lg, ft

    def forward(self, x, only_fc=False, only_feat=False, **kwargs):
        if only_fc:   # hubert.py:31-33: the classifier alone on pooled features, off the train-step path
            return self.classifier(x)
        if torch.is_grad_enabled() and self.training:
            raise RuntimeError("the native HuBERT is driven by the SSL step's eager backward (forward_native / backward_native); "
                               "a plain autograd forward is not provided — call under torch.no_grad() for inference")
        logits, feat = self._infer(x)
        if only_feat:
            return feat
        return {"logits": logits, "feat": feat}

    def extract(self, x):
        return self._infer(x)[1]

    def group_matcher(self, coarse=False, prefix=""):
        return dict(stem=r"^{}model.feature_projection|^{}model.feature_extractor|^{}model.encoder.pos_conv_embed".format(prefix, prefix, prefix),
                    blocks=r"^{}model.encoder.layers.(\d+)".format(prefix))

    def no_weight_decay(self):
        return []


def _load_pretrained(model, path):
    sd = torch.load(path, map_location="cpu")
    sd = sd.get("model", sd.get("state_dict", sd))
    own = model.state_dict()
    fixed = {}
    for k, v in sd.items():
        if k.startswith("module."):
            k = k[7:]
        # pre-parametrization checkpoints name the weight-norm tensors weight_g / weight_v
        k = k.replace("pos_conv_embed.conv.weight_g", "pos_conv_embed.conv.parametrizations.weight.original0")
        k = k.replace("pos_conv_embed.conv.weight_v", "pos_conv_embed.conv.parametrizations.weight.original1")
        if k in own:
            fixed[k] = v
        elif "model." + k in own:
            fixed["model." + k] = v
    print(model.load_state_dict(fixed, strict=False))
    model.mark_weights_updated()
    return model


def hubert_base(pretrained=False, pretrained_path=None, **kwargs):
    model = ClassificationHubert(name="facebook/hubert-base-ls960", **kwargs)
    if pretrained and pretrained_path and os.path.isfile(str(pretrained_path)):
        model = _load_pretrained(model, pretrained_path)
    return model
