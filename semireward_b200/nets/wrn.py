"""B200-native WideResNet behind the reference's net-builder interface (semilearn/nets/wrn/wrn.py:30-173).

  * builders `wrn_28_2` / `wrn_28_8` (wrn.py:162-173): f(pretrained=False, pretrained_path=None, **kw) -> nn.Module
  * module contract (wrn.py:118-160): forward(x, only_fc=False, only_feat=False) -> {'logits', 'feat'}, extract(), group_matcher(),
    no_weight_decay(), num_features = channels = 64 * widen_factor
  * identical module tree, hence identical state_dict keys / shapes / order INCLUDING the BatchNorm buffers (running_mean, running_var,
    num_batches_tracked): 81 parameters + 75 buffers for WRN-28-2.  Initialisation as wrn.py:108-116.
All arithmetic runs in libsrw_b200.so (srw_wrn_forward / srw_wrn_backward); the nn.Conv2d / nn.BatchNorm2d / nn.Linear children are
parameter and buffer holders and are never called.  There is no PyTorch fallback.

BatchNorm couples the rows of a launch, so (unlike the LayerNorm backbones) the K sampling passes of stage 2 cannot be merged into one
forward; but the network has no dropout (dropRate 0 in every config), so those passes recompute the identical tensors and only advance
the running statistics — `stat_repeats` of srw_wrn_forward does exactly that (include/srw.h).  Data parallel = SyncBatchNorm
(core/utils/misc.py:54 converts every BatchNorm2d under DDP): when the module carries a data-parallel group (send_model_cuda), the engine
hands every layer's per-channel sums to `_sync_sums`, which all-reduces them over the group on the current stream; those calls run
eagerly (a host-enqueued collective between kernels cannot be replayed from a CUDA graph)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from .. import _lib as L
from ._native import NativeBackbone
from .utils import load_checkpoint

momentum = 0.001


def _conv(cin, cout, k, stride, bias=False):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=bias)


class BasicBlock(nn.Module):
    """Parameter / buffer holder of one pre-activation block.  Children carry the reference's names in its registration order
    (bn1, conv1, bn2, conv2[, convShortcut]; wrn.py:30-44), so `state_dict()` keys and `named_parameters()` order are the reference's;
    the activation modules have no state and are not materialised."""

    def __init__(self, in_planes, out_planes, stride, activate_before_residual=False):
        super().__init__()
        self.equalInOut = in_planes == out_planes
        self.activate_before_residual = activate_before_residual
        for name, mod in (("bn1", nn.BatchNorm2d(in_planes, momentum=momentum)), ("conv1", _conv(in_planes, out_planes, 3, stride)),
                          ("bn2", nn.BatchNorm2d(out_planes, momentum=momentum)), ("conv2", _conv(out_planes, out_planes, 3, 1))):
            self.add_module(name, mod)
        self.convShortcut = None if self.equalInOut else _conv(in_planes, out_planes, 1, stride)


class NetworkBlock(nn.Module):
    """`nb_layers` blocks under `.layer` (an nn.Sequential, wrn.py:57-66): only the first changes width / resolution."""

    def __init__(self, nb_layers, in_planes, out_planes, stride, activate_before_residual=False):
        super().__init__()
        widths = [in_planes] + [out_planes] * (int(nb_layers) - 1)
        self.layer = nn.Sequential(*[BasicBlock(w, out_planes, stride if i == 0 else 1, activate_before_residual) for i, w in enumerate(widths)])


class WideResNet(NativeBackbone, nn.Module):
    def __init__(self, first_stride, num_classes, depth=28, widen_factor=2, drop_rate=0.0, img_size=32, leaky_slope=0.1, **kwargs):
        super().__init__()
        if first_stride != 1:
            raise NotImplementedError("native WideResNet: first_stride must be 1 (wrn_28_2 / wrn_28_8)")
        if drop_rate != 0.0:
            raise NotImplementedError("native WideResNet: drop_rate > 0 is not built (0 in every shipped config)")
        channels = [16, 16 * widen_factor, 32 * widen_factor, 64 * widen_factor]
        assert (depth - 4) % 6 == 0
        n = (depth - 4) // 6
        self.conv1 = _conv(3, channels[0], 3, 1, bias=True)
        for b, stride in enumerate((first_stride, 2, 2)):
            self.add_module(f"block{b + 1}", NetworkBlock(n, channels[b], channels[b + 1], stride, activate_before_residual=(b == 0)))
        self.bn1 = nn.BatchNorm2d(channels[3], momentum=momentum, eps=0.001)      # wrn.py:97: the one BatchNorm with eps 1e-3
        self.classifier = nn.Linear(channels[3], num_classes)
        self.channels = channels[3]
        self.num_features = channels[3]
        for m in self.modules():   # wrn.py:108-116
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="leaky_relu")
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight.data)
                m.bias.data.zero_()
        self.gemm_impl = L.GEMM_TCGEN05
        self.depth, self.widen, self.img_size = depth, widen_factor, img_size
        self._cfg = L.WrnConfig(num_classes=num_classes, depth=depth, widen=widen_factor, img_size=img_size, bn_momentum=momentum, slope=leaky_slope)   # 0.1 in the reference (wrn.py:34,38,98); a test knob
        self._planes = self._planes_key = None
        self._init_native()
        self.stat_repeats_next = 0     # set by the SSL step before the forward of a stage-2 step (K identical sampling passes)
        self._bn_key = self._bn_arrays = None

    # -- native plumbing ------------------------------------------------------------------------
    def _blocks(self):
        return list(self.block1.layer) + list(self.block2.layer) + list(self.block3.layer)

    def _ordered_params(self):
        return [p for _, p in self.named_parameters()]   # the module tree registers them in the engine's order (include/srw.h)

    def _grad_params(self):
        """The bn1 of the first layer of block2 / block3 gets no gradient (its output is dropped, wrn.py:46-51): like the reference,
        those two `p.grad` stay None and the optimizer skips them."""
        dead = set()
        for b in (self.block2.layer[0], self.block3.layer[0]):
            dead |= {id(b.bn1.weight), id(b.bn1.bias)}
        return [p for p in self._ordered_params() if id(p) not in dead]

    def _bn_modules(self):
        out = []
        for b in self._blocks():
            out += [b.bn1, b.bn2]
        return out + [self.bn1]

    def _bn_pointers(self):
        bns = self._bn_modules()
        key = tuple(m.running_mean.data_ptr() for m in bns)
        if key != self._bn_key:
            self._bn_arrays = (L.ptr_array([m.running_mean for m in bns]), L.ptr_array([m.running_var for m in bns]),
                               L.ptr_array([m.num_batches_tracked for m in bns]))
            self._bn_key = key
        return self._bn_arrays

    def _weight_planes(self):
        ps = self._ordered_params()
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._planes is None or key != self._planes_key:
            lib = L.load()
            dev = ps[0].device
            if self._planes is None or self._planes.device != dev:
                self._planes = torch.empty(lib.srw_wrn_weight_planes_bytes(C.byref(self._cfg)), dtype=torch.uint8, device=dev)
            L.check(lib.srw_wrn_prepare_weights(C.byref(self._cfg), L.ptr_array([p.detach() for p in ps]), self._planes.data_ptr(), L.stream_ptr()),
                    "srw_wrn_prepare_weights")
            self._planes_key = key
        return self._planes

    def weight_plane_slot(self, idx):
        return None     # re-laid-out convolution operands: rebuilt by srw_wrn_prepare_weights after every step (12 MB)

    def mark_weights_updated(self, planes_fresh: bool = False):
        self._planes_key = None

    def stochastic(self):
        return False

    def draw_streams(self, num_passes, nl, nu, device):
        return 0

    def streams_for(self, draws, pieces, nl, nu, device):
        return None

    def concat_inputs(self, parts, device):
        S = sum(int(p.shape[0]) for p in parts)
        buf = self._buf("x", (S,) + tuple(parts[0].shape[1:]), device)
        torch.cat([p.to(torch.float32) for p in parts], out=buf)
        return buf

    # -- engine calls ---------------------------------------------------------------------------
    @torch.no_grad()
    def forward_native(self, x, grad_batch=0, drop_scale=None):
        """One autograd-free forward through srw_wrn_forward -> (logits, feat, handle).  BatchNorm runs in the module's mode
        (train: batch statistics of ALL rows of x + running-statistics update)."""
        if not x.is_cuda:
            raise RuntimeError("semireward_b200 WideResNet runs on CUDA (sm_100a) only; there is no CPU path")
        lib, cfg, dev = L.load(), self._cfg, x.device
        S = x.shape[0]
        if tuple(x.shape[1:]) != (3, cfg.img_size, cfg.img_size):
            raise ValueError(f"native WideResNet was built for 3 x {cfg.img_size} x {cfg.img_size} images, got {tuple(x.shape[1:])}")
        x = x.contiguous()
        params, pa = self._native_params()
        rm, rv, nbt = self._bn_pointers()
        wbytes = lib.srw_wrn_workspace_bytes(C.byref(cfg), S)
        if wbytes < 0:
            L.check(-2, "srw_wrn_workspace_bytes")
        ws = self._acquire_ws(wbytes, dev)
        lo, fe = self._buf("logits", (S, cfg.num_classes), dev), self._buf("feat", (S, self.channels), dev)
        rep, self.stat_repeats_next = (int(self.stat_repeats_next) if self.training else 0), 0
        sync = self._sync_args(dev) if self.training else {}
        a = L.WrnFwdArgs(cfg=C.pointer(cfg), params=pa, bn_running_mean=rm, bn_running_var=rv, bn_num_batches_tracked=nbt,
                         weight_planes=self._weight_planes().data_ptr(), x=x.data_ptr(), batch=S, training=int(self.training), stat_repeats=rep,
                         logits=lo.data_ptr(), feat=fe.data_ptr(), workspace=ws.data_ptr(), workspace_bytes=wbytes, gemm_impl=self.gemm_impl, **sync)
        L.check(lib.srw_wrn_forward(C.byref(a), L.stream_ptr()), "srw_wrn_forward")
        handle = dict(ws=ws, wbytes=wbytes, x=x, B=S, grad_batch=grad_batch, sync=sync)
        if grad_batch == 0:
            self.release_pass(handle)
        return lo.clone(), fe.clone(), handle

    def dlogits_buffer(self, grad_batch, device):
        return self._buf("dlogits", (grad_batch, self._cfg.num_classes), device)

    # -- SyncBatchNorm ------------------------------------------------------------------------------
    def _sync_args(self, dev):
        """Arguments that switch the engine's BatchNorms to statistics over all ranks of the data-parallel group, or {} (single rank)."""
        group = getattr(self, "_dp_group", None)
        if group is None:
            return {}
        import torch.distributed as dist
        world = dist.get_world_size(group)
        if world == 1:
            return {}
        if getattr(self, "_sync_cb", None) is None or self._sync_group is not group:
            buf = self._buf("syncbn_sums", (2 * self.channels,), dev)

            def _sync_sums(ctx, ptr, count):   # called by the engine on this thread, between two kernels of the current stream
                try:
                    off = (int(ptr) - buf.data_ptr()) // 4
                    dist.all_reduce(buf[off:off + count], op=dist.ReduceOp.SUM, group=group)
                    return 0
                except Exception:   # noqa: BLE001 — an exception must not unwind through the C frame
                    import traceback
                    traceback.print_exc()
                    return -1
            self._sync_cb, self._sync_buf, self._sync_group = L.ALLREDUCE_SUM_FN(_sync_sums), buf, group
        return dict(sync_fn=C.cast(self._sync_cb, C.c_void_p), sync_ctx=None, sync_buf=self._sync_buf.data_ptr(), world_size=world)

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
            df = self._buf("dfeat", (Sg, self.channels), dev)
            df.copy_(dfeat)
        a = L.WrnBwdArgs(cfg=C.pointer(cfg), params=pa, weight_planes=self._weight_planes().data_ptr(), batch=handle["B"], grad_rows=Sg, dlogits=dl.data_ptr(),
                         dfeat=L.ptr(df), grads=self._ga, accumulate_grads=int(bool(accumulate)), workspace=handle["ws"].data_ptr(),
                         workspace_bytes=handle["wbytes"], gemm_impl=self.gemm_impl, **handle["sync"])
        L.check(lib.srw_wrn_backward(C.byref(a), L.stream_ptr()), "srw_wrn_backward")
        self._pending_reduce = []
        self.release_pass(handle)
        return self._flat_grads, self._grad_views

    # -- reference interface --------------------------------------------------------------------
    @torch.no_grad()
    def _infer(self, x):
        lg, ft, _ = self.forward_native(self.concat_inputs([x], x.device), grad_batch=0)
        return lg, ft

    def forward(self, x, only_fc=False, only_feat=False, **kwargs):
        if only_fc:
            return self.classifier(x)
        if torch.is_grad_enabled() and self.training:
            raise RuntimeError("the native WideResNet is driven by the SSL step's eager backward (forward_native / backward_native); "
                               "a plain autograd forward is not provided — call under torch.no_grad() for inference")
        logits, feat = self._infer(x)
        if only_feat:
            return feat
        return {"logits": logits, "feat": feat}

    def extract(self, x):
        raise NotImplementedError("native WideResNet: extract() (the un-pooled feature map, wrn.py:138-146) is not produced by the fused engine")

    def group_matcher(self, coarse=False, prefix=""):
        return dict(stem=r"^{}conv1".format(prefix), blocks=r"^{}block(\d+)".format(prefix) if coarse else r"^{}block(\d+)\.layer.(\d+)".format(prefix))

    def no_weight_decay(self):
        return [n for n, _ in self.named_parameters() if "bn" in n or "bias" in n]


def wrn_28_2(pretrained=False, pretrained_path=None, **kwargs):
    model = WideResNet(first_stride=1, depth=28, widen_factor=2, **kwargs)
    if pretrained:
        model = load_checkpoint(model, pretrained_path)
        model.mark_weights_updated()
    return model


def wrn_28_8(pretrained=False, pretrained_path=None, **kwargs):
    model = WideResNet(first_stride=1, depth=28, widen_factor=8, **kwargs)
    if pretrained:
        model = load_checkpoint(model, pretrained_path)
        model.mark_weights_updated()
    return model
