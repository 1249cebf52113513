"""B200-native ClassificationBert behind the reference's net-builder interface (semilearn/nets/bert/bert.py:9-75).

  * builders `bert_base_uncased` / `bert_base_cased` (bert.py:68-75): f(pretrained=True, pretrained_path=None, **kw) -> nn.Module
  * module contract (bert.py:22-66): forward(x = {'input_ids', 'attention_mask'}, only_fc=False, only_feat=False) ->
    {'logits', 'feat'}, extract(), group_matcher(), no_weight_decay(), num_features = 768
  * identical state_dict keys / shapes / registration order as the reference's module (Hugging Face `BertModel` under `bert.`,
    pooler included, then `classifier.0`, `classifier.2`): 205 tensors for bert-base.
The reference builds the encoder with `BertModel.from_pretrained(name)` (a hub download); offline, and for random-init synthetic
runs, the builder initialises like `BertModel(BertConfig())` (normal(0, 0.02) matrices, zero biases, unit
LayerNorms, zero padding row) and loads `pretrained_path` when it names a local state-dict file (HF key names, with or without the
`bert.` prefix).  All arithmetic runs in libsrw_b200.so (srw_bert_forward / srw_bert_backward); the nn.Linear / nn.Embedding /
nn.LayerNorm children are parameter holders and are never called.  There is no PyTorch fallback."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L
from ._native import NativeBackbone


def _mix32(x: int) -> int:
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    return x ^ (x >> 16)


def call_key(seed: int, call: int) -> int:
    """Dropout stream key of one backbone call (include/srw.h: srw_dropout.seq_key): the reference draws fresh nn.Dropout masks in
    every one of its model calls (three per step with `use_cat: False`, three more per sampling pass)."""
    return _mix32(seed * 0x9E3779B9 + call * 0x85EBCA6B + 0x165667B1)


class _SelfAttention(nn.Module):
    def __init__(self, H):
        super().__init__()
        self.query, self.key, self.value = nn.Linear(H, H), nn.Linear(H, H), nn.Linear(H, H)


class _DenseLN(nn.Module):
    def __init__(self, i, o, eps):
        super().__init__()
        self.dense = nn.Linear(i, o)
        self.LayerNorm = nn.LayerNorm(o, eps=eps)


class _Dense(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.dense = nn.Linear(i, o)


class _Attention(nn.Module):
    def __init__(self, H, eps):
        super().__init__()
        self.self = _SelfAttention(H)
        self.output = _DenseLN(H, H, eps)


class _Layer(nn.Module):
    def __init__(self, H, I, eps):
        super().__init__()
        self.attention = _Attention(H, eps)
        self.intermediate = _Dense(H, I)
        self.output = _DenseLN(I, H, eps)


class _Encoder(nn.Module):
    def __init__(self, H, I, n, eps):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(H, I, eps) for _ in range(n)])


class _Embeddings(nn.Module):
    def __init__(self, vocab, max_pos, type_vocab, H, eps):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, H, padding_idx=0)
        self.position_embeddings = nn.Embedding(max_pos, H)
        self.token_type_embeddings = nn.Embedding(type_vocab, H)
        self.LayerNorm = nn.LayerNorm(H, eps=eps)


class _BertModel(nn.Module):  # parameter holder with Hugging Face BertModel's module tree
    def __init__(self, vocab, max_pos, type_vocab, H, I, n, eps):
        super().__init__()
        self.embeddings = _Embeddings(vocab, max_pos, type_vocab, H, eps)
        self.encoder = _Encoder(H, I, n, eps)
        self.pooler = _Dense(H, H)


class ClassificationBert(NativeBackbone, nn.Module):
    def __init__(self, name="bert-base-uncased", num_classes=2, vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12, hidden_dropout_prob=0.1,
                 attention_probs_dropout_prob=0.1, pooled_dropout=0.1, initializer_range=0.02):
        super().__init__()
        if name == "bert-base-cased" and vocab_size == 30522:
            vocab_size = 28996
        H = hidden_size
        self.bert = _BertModel(vocab_size, max_position_embeddings, type_vocab_size, H, intermediate_size, num_hidden_layers, layer_norm_eps)
        self.dropout = nn.Dropout(p=pooled_dropout, inplace=False)   # holder of p (bert.py:14); the mask is drawn inside the engine
        self.num_features = H
        self.classifier = nn.Sequential(nn.Linear(H, H), nn.GELU(), nn.Linear(H, num_classes))
        with torch.no_grad():   # BertPreTrainedModel._init_weights for the encoder; the classifier keeps nn.Linear's default like the reference
            for m in self.bert.modules():
                if isinstance(m, nn.Linear):
                    m.weight.normal_(0.0, initializer_range)
                    m.bias.zero_()
                elif isinstance(m, nn.Embedding):
                    m.weight.normal_(0.0, initializer_range)
                    if m.padding_idx is not None:
                        m.weight[m.padding_idx].zero_()
        self.gemm_impl = L.GEMM_TCGEN05
        self._cfg = L.BertConfig(vocab_size=vocab_size, max_position=max_position_embeddings, type_vocab=type_vocab_size, hidden=H,
                                 layers=num_hidden_layers, heads=num_attention_heads, intermediate=intermediate_size, num_classes=num_classes,
                                 ln_eps=layer_norm_eps, p_hidden=hidden_dropout_prob, p_attn=attention_probs_dropout_prob, p_pooled=pooled_dropout)
        self._planes = self._planes_key = None
        self._init_native()
        self.dropout_seed = 0          # seed of the counter-based dropout streams (call_key)
        self._calls = 0                # backbone calls made so far: every call of the reference draws fresh masks
        self.depth = num_hidden_layers

    # -- native plumbing ------------------------------------------------------------------------
    def _ordered_params(self):
        """Engine order == state_dict order (include/srw.h); the pooler's two tensors are present in the list but never read."""
        e = self.bert.embeddings
        ps = [e.word_embeddings.weight, e.position_embeddings.weight, e.token_type_embeddings.weight, e.LayerNorm.weight, e.LayerNorm.bias]
        for ly in self.bert.encoder.layer:
            a = ly.attention
            ps += [a.self.query.weight, a.self.query.bias, a.self.key.weight, a.self.key.bias, a.self.value.weight, a.self.value.bias,
                   a.output.dense.weight, a.output.dense.bias, a.output.LayerNorm.weight, a.output.LayerNorm.bias,
                   ly.intermediate.dense.weight, ly.intermediate.dense.bias, ly.output.dense.weight, ly.output.dense.bias,
                   ly.output.LayerNorm.weight, ly.output.LayerNorm.bias]
        ps += [self.bert.pooler.dense.weight, self.bert.pooler.dense.bias, self.classifier[0].weight, self.classifier[0].bias,
               self.classifier[2].weight, self.classifier[2].bias]
        return ps

    def _grad_params(self):
        """Flat gradient buffer order: the engine needs d(query | key | value).weight contiguous and likewise their biases (one
        packed projection GEMM, include/srw.h).  The pooler has no gradient (its output is unused, bert.py:35), as in the reference,
        where those two `p.grad` stay None and AdamW skips them."""
        e = self.bert.embeddings
        ps = [e.word_embeddings.weight, e.position_embeddings.weight, e.token_type_embeddings.weight, e.LayerNorm.weight, e.LayerNorm.bias]
        for ly in self.bert.encoder.layer:
            a = ly.attention
            ps += [a.self.query.weight, a.self.key.weight, a.self.value.weight, a.self.query.bias, a.self.key.bias, a.self.value.bias,
                   a.output.dense.weight, a.output.dense.bias, a.output.LayerNorm.weight, a.output.LayerNorm.bias,
                   ly.intermediate.dense.weight, ly.intermediate.dense.bias, ly.output.dense.weight, ly.output.dense.bias,
                   ly.output.LayerNorm.weight, ly.output.LayerNorm.bias]
        ps += [self.classifier[0].weight, self.classifier[0].bias, self.classifier[2].weight, self.classifier[2].bias]
        return ps

    def _matrix_params(self):
        ps = []
        for ly in self.bert.encoder.layer:
            a = ly.attention
            ps += [a.self.query.weight, a.self.key.weight, a.self.value.weight, a.output.dense.weight, ly.intermediate.dense.weight, ly.output.dense.weight]
        return ps

    def _weight_planes(self):
        mats = self._matrix_params()
        key = tuple((p.data_ptr(), p._version) for p in mats)
        if self._planes is None or key != self._planes_key:
            lib = L.load()
            dev = mats[0].device
            if self._planes is None or self._planes.device != dev:
                self._planes = torch.empty(lib.srw_bert_weight_planes_bytes(C.byref(self._cfg)), dtype=torch.uint8, device=dev)
            params = [p.detach() for p in self._ordered_params()]
            L.check(lib.srw_bert_prepare_weights(C.byref(self._cfg), L.ptr_array(params), self._planes.data_ptr(), L.stream_ptr()),
                    "srw_bert_prepare_weights")
            self._planes_key = key
        return self._planes

    def weight_plane_slot(self, idx):
        """(byte offset, cols, ldp, plane stride) of parameter `idx` (engine order) inside the weight-plane cache, or None."""
        off, cols, ldp, ps_ = L.i64(), L.i32(), L.i32(), L.i64()
        if L.load().srw_bert_weight_plane_slot(C.byref(self._cfg), idx, C.byref(off), C.byref(cols), C.byref(ldp), C.byref(ps_)) == 0:
            return off.value, cols.value, ldp.value, ps_.value
        return None

    def mark_weights_updated(self, planes_fresh: bool = False):
        if planes_fresh and self._planes is not None:
            self._planes_key = tuple((p.data_ptr(), p._version) for p in self._matrix_params())
        else:
            self._planes_key = None

    def stochastic(self):
        c = self._cfg
        return self.training and max(c.p_hidden, c.p_attn, c.p_pooled) > 0.0

    # -- dropout streams: which (stream key, row) every sequence of a launch uses ------------------------------------------------
    def draw_streams(self, num_passes, nl, nu, device):
        """Reserve the dropout streams of `num_passes` passes of three calls each (labelled, strong, weak: the reference's call order,
        srsoftmatch.py:119-130) -> first call index."""
        first = self._calls
        self._calls += 3 * num_passes
        return first

    def streams_for(self, draws, pieces, nl, nu, device):
        """pieces: [(pass, 'lb' | 's' | 'w'), ...] in launch row order -> per-sequence (key int32-bits [S], row int32 [S]) or None."""
        if not self.stochastic():
            return None
        keys, rows = [], []
        for ps_, part in pieces:
            n = nl if part == "lb" else nu
            k = call_key(self.dropout_seed, draws + 3 * ps_ + {"lb": 0, "s": 1, "w": 2}[part])
            keys += [k] * n
            rows += list(range(n))
        kt = torch.from_numpy(np.asarray(keys, dtype=np.uint32).view(np.int32).copy())
        rt = torch.tensor(rows, dtype=torch.int32)
        return kt, rt

    def concat_inputs(self, parts, device):
        """parts: dicts {'input_ids', 'attention_mask'} -> persistent (ids [S, L], mask [S, L], pool_len [S]) on `device`.  Calls of
        different padded length are right-padded to the longest (padding id 0, mask 0); `pool_len` keeps every call's own
        length, over which the reference's mean pool runs (bert.py:36-37)."""
        Lmax = max(p["input_ids"].shape[1] for p in parts)
        S = sum(p["input_ids"].shape[0] for p in parts)
        ids, am, pl = self._buf("ids", (S, Lmax), device, torch.long), self._buf("mask", (S, Lmax), device, torch.long), self._buf("pool_len", (S,), device, torch.int32)
        same = all(p["input_ids"].shape[1] == Lmax for p in parts)
        if same:
            torch.cat([p["input_ids"] for p in parts], out=ids)
            if all(p.get("attention_mask") is not None for p in parts):
                torch.cat([p["attention_mask"] for p in parts], out=am)
            else:
                am.fill_(1)
            pl.fill_(Lmax)
        else:
            ids.zero_(); am.zero_()
            r = 0
            for p in parts:
                n, ln = p["input_ids"].shape
                ids[r:r + n, :ln].copy_(p["input_ids"])
                am[r:r + n, :ln].copy_(p["attention_mask"] if p.get("attention_mask") is not None else torch.ones_like(p["input_ids"]))
                pl[r:r + n].fill_(ln)
                r += n
        return ids, am, pl

    # -- engine calls ---------------------------------------------------------------------------
    @torch.no_grad()
    def forward_native(self, x, grad_batch=0, drop_scale=None):
        """One autograd-free forward through srw_bert_forward -> (logits, feat, handle).  x = (ids, mask, pool_len) from
        concat_inputs(); drop_scale = streams_for(...) or None (no dropout)."""
        ids, am, pl = x
        if not ids.is_cuda:
            raise RuntimeError("semireward_b200 BERT runs on CUDA (sm_100a) only; there is no CPU path")
        lib, cfg, dev = L.load(), self._cfg, ids.device
        S, Lq = ids.shape
        if Lq < 16:
            raise ValueError("sequence length must be >= 16 (pad the batch)")
        params, pa = self._native_params()
        wbytes = lib.srw_bert_workspace_bytes(C.byref(cfg), S, Lq, grad_batch)
        if wbytes < 0:
            L.check(-2, "srw_bert_workspace_bytes")
        ws = self._acquire_ws(wbytes, dev)
        kt = rt = None
        if drop_scale is not None:
            kt, rt = self._buf(("dkey", ws.data_ptr()), (S,), dev, torch.int32), self._buf(("drow", ws.data_ptr()), (S,), dev, torch.int32)
            kt.copy_(drop_scale[0], non_blocking=True)
            rt.copy_(drop_scale[1], non_blocking=True)
        lo, fe = self._buf("logits", (S, cfg.num_classes), dev), self._buf("feat", (S, cfg.hidden), dev)
        a = L.BertFwdArgs(cfg=C.pointer(cfg), params=pa, weight_planes=self._weight_planes().data_ptr(), input_ids=ids.data_ptr(),
                          attention_mask=am.data_ptr(), batch=S, seq_len=Lq, grad_batch=grad_batch, drop_seq_key=L.ptr(kt), drop_seq_row=L.ptr(rt),
                          pool_len=L.ptr(pl), logits=lo.data_ptr(), feat=fe.data_ptr(), workspace=ws.data_ptr(), workspace_bytes=wbytes,
                          gemm_impl=self.gemm_impl)
        L.check(lib.srw_bert_forward(C.byref(a), L.stream_ptr()), "srw_bert_forward")
        handle = dict(ws=ws, wbytes=wbytes, x=x, keys=kt, rows=rt, B=S, L=Lq, grad_batch=grad_batch)
        if grad_batch == 0:
            self.release_pass(handle)
        return lo.clone(), fe.clone(), handle

    def dlogits_buffer(self, grad_batch, device):
        return self._buf("dlogits", (grad_batch, self._cfg.num_classes), device)

    def _layer_offset(self, lo):
        """Offset (elements) of layer `lo`'s first gradient inside the flat buffer (order of _grad_params)."""
        gps = self._grad_params()
        return sum(p.numel() for p in gps[:5 + 16 * lo])

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
        ids, am, pl = handle["x"]
        a = L.BertBwdArgs(cfg=C.pointer(cfg), params=pa, weight_planes=self._weight_planes().data_ptr(), input_ids=ids.data_ptr(),
                          attention_mask=am.data_ptr(), batch=handle["B"], seq_len=handle["L"], grad_batch=Sg, drop_seq_key=L.ptr(handle["keys"]),
                          drop_seq_row=L.ptr(handle["rows"]), pool_len=L.ptr(pl), dlogits=dl.data_ptr(), dfeat=L.ptr(df), grads=self._ga,
                          accumulate_grads=int(bool(accumulate)), workspace=handle["ws"].data_ptr(), workspace_bytes=handle["wbytes"],
                          gemm_impl=self.gemm_impl, layer_lo=-1, layer_hi=-1)
        group = getattr(self, "_dp_group", None)
        self._pending_reduce = []
        bounds = self._dp_bounds(cfg.layers) if (group is not None and final) else []
        if bounds:   # DDP-style overlap: the gradients of a finished layer range are a contiguous tail piece of the flat buffer
            hi, end = cfg.layers - 1, self._flat_grads.numel()
            for lo in bounds + [0]:
                a.layer_lo, a.layer_hi = lo, hi
                L.check(lib.srw_bert_backward(C.byref(a), L.stream_ptr()), "srw_bert_backward")
                off = self._layer_offset(lo) if lo > 0 else 0
                self._pending_reduce.append(self._allreduce_async(self._flat_grads[off:end], group))
                hi, end = lo - 1, off
        else:
            L.check(lib.srw_bert_backward(C.byref(a), L.stream_ptr()), "srw_bert_backward")
        self.release_pass(handle)
        return self._flat_grads, self._grad_views

    # -- reference interface --------------------------------------------------------------------
    @torch.no_grad()
    def _infer(self, x):
        dev = x["input_ids"].device
        n = x["input_ids"].shape[0]
        inp = self.concat_inputs([x], dev)
        spec = None
        if self.stochastic():   # a train-mode call outside the SSL step still draws dropout, like the reference's module would
            d = self.draw_streams(1, n, 0, dev)
            spec = self.streams_for(d, [(0, "lb")], n, 0, dev)
        lg, ft, _ = self.forward_native(inp, grad_batch=0, drop_scale=spec)
        return lg, ft

    def forward(self, x, only_fc=False, only_feat=False, return_embed=False, **kwargs):
        if only_fc:   # bert.py:30-32: the classifier alone on pooled features, off the train-step path
            return self.classifier(x)
        if return_embed:
            raise NotImplementedError("return_embed (word embeddings for VAT) is not produced by the fused engine")
        if torch.is_grad_enabled() and self.training:
            raise RuntimeError("the native BERT is driven by the SSL step's eager backward (forward_native / backward_native); "
                               "a plain autograd forward is not provided — call under torch.no_grad() for inference")
        logits, feat = self._infer(x)
        if only_feat:
            return feat
        return {"logits": logits, "feat": feat}

    def extract(self, x):
        return self._infer(x)[1]

    def group_matcher(self, coarse=False, prefix=""):
        return dict(stem=r"^{}bert.embeddings".format(prefix), blocks=r"^{}bert.encoder.layer.(\d+)".format(prefix))

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
        if k in own:
            fixed[k] = v
        elif "bert." + k in own:
            fixed["bert." + k] = v
    print(model.load_state_dict(fixed, strict=False))
    model.mark_weights_updated()
    return model


def _build(name, pretrained, pretrained_path, kwargs):
    model = ClassificationBert(name=name, **kwargs)
    if pretrained and pretrained_path and os.path.isfile(str(pretrained_path)):
        model = _load_pretrained(model, pretrained_path)
    return model


def bert_base_cased(pretrained=True, pretrained_path=None, **kwargs):
    return _build("bert-base-cased", pretrained, pretrained_path, kwargs)


def bert_base_uncased(pretrained=True, pretrained_path=None, **kwargs):
    return _build("bert-base-uncased", pretrained, pretrained_path, kwargs)
