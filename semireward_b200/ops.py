"""Thin Python wrappers over the C ABI (include/srw.h) for single kernels.  Device memory comes from torch tensors
(plumbing only); all arithmetic happens inside libsrw_b200.so.  Used by the parity tests and by the host-side
plugin for the pieces that are not inside the native ViT engine."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def _lib():
    return L.load()


def _s():
    return L.stream_ptr()


class Planes:
    """Split-bf16 operand: tensor `t` of shape [2, rows, ld] (hi plane, lo plane), logical width `cols`."""

    def __init__(self, t: torch.Tensor, rows: int, cols: int):
        assert t.dtype == torch.bfloat16 and t.dim() == 3 and t.shape[0] == 2
        self.t, self.rows, self.cols = t, rows, cols

    @property
    def ld(self) -> int:
        return self.t.stride(1)

    @property
    def plane_stride(self) -> int:
        return self.t.stride(0)

    def ptr(self) -> int:
        return self.t.data_ptr()

    def to_f32(self) -> torch.Tensor:
        return (self.t[0].float() + self.t[1].float())[:, :self.cols]


def empty_planes(rows: int, cols: int, device="cuda") -> Planes:
    return Planes(torch.empty(2, rows, cols, dtype=torch.bfloat16, device=device), rows, cols)


def split_planes(x: torch.Tensor, transposed: bool = False, row_scale: torch.Tensor | None = None, rows_per_scale: int = 1):
    """x [rows, cols] fp32 -> Planes (and the transposed Planes [cols, rows] when transposed=True)."""
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    p = empty_planes(rows, cols, x.device)
    pt = empty_planes(cols, rows, x.device) if transposed else None
    a = L.SplitArgs(x=x.data_ptr(), ldx=x.stride(0), rows=rows, cols=cols, row_scale=L.ptr(row_scale), rows_per_scale=rows_per_scale,
                    planes=p.ptr(), ldp=p.ld, plane_stride=p.plane_stride,
                    planes_t=pt.ptr() if pt else None, ldpt=pt.ld if pt else 0, plane_stride_t=pt.plane_stride if pt else 0)
    L.check(_lib().srw_split_planes(C.byref(a), _s()), "srw_split_planes")
    return (p, pt) if transposed else p


def gemm(a: Planes, b: Planes, M: int, N: int, K: int, a_mn: bool = False, b_mn: bool = False, epilogue: int = L.EPI_F32,
         bias=None, resid=None, row_scale=None, rows_per_scale: int = 1, aux=None, out_f32=None, out_planes: Planes | None = None,
         split_k: int = 1, workspace=None, impl: int = L.GEMM_TCGEN05, drop: L.Dropout | None = None, drop_rows_per_seq: int = 1):
    """D[M,N] = A[M,K] * B[N,K]^T (+ epilogue).  Returns (out_f32, out_planes) as applicable."""
    dev = a.t.device
    if epilogue in (L.EPI_F32, L.EPI_GELU, L.EPI_RESID) and out_f32 is None:
        out_f32 = torch.empty(M, N, dtype=torch.float32, device=dev)
    if epilogue in (L.EPI_PLANES, L.EPI_GELU, L.EPI_DGELU) and out_planes is None:
        out_planes = empty_planes(M, N, dev)
    if epilogue == L.EPI_SPLITK and workspace is None:
        workspace = torch.empty(split_k, M, N, dtype=torch.float32, device=dev)
    g = L.GemmArgs(M=M, N=N, K=K, a=a.ptr(), lda=a.ld, a_plane_stride=a.plane_stride, a_mn_major=int(a_mn),
                   b=b.ptr(), ldb=b.ld, b_plane_stride=b.plane_stride, b_mn_major=int(b_mn), epilogue=epilogue,
                   bias=L.ptr(bias), resid=L.ptr(resid), ldr=resid.stride(0) if resid is not None else 0,
                   row_scale=L.ptr(row_scale), rows_per_scale=rows_per_scale, aux=L.ptr(aux),
                   ldaux=aux.stride(0) if aux is not None else 0, out_f32=L.ptr(out_f32),
                   ldo=out_f32.stride(0) if out_f32 is not None else 0,
                   out_planes=out_planes.ptr() if out_planes else None, ldp=out_planes.ld if out_planes else 0,
                   out_plane_stride=out_planes.plane_stride if out_planes else 0, split_k=split_k,
                   workspace=L.ptr(workspace), impl=impl, drop=drop if drop is not None else dropout_spec(),
                   drop_rows_per_seq=drop_rows_per_seq)
    L.check(_lib().srw_gemm(C.byref(g), _s()), "srw_gemm")
    if epilogue == L.EPI_SPLITK:
        return workspace
    return out_f32, out_planes


def splitk_reduce(workspace: torch.Tensor, out: torch.Tensor, accumulate: bool = False):
    split, M, N = workspace.shape
    a = L.SplitKReduceArgs(workspace=workspace.data_ptr(), split_k=split, M=M, N=N, out=out.data_ptr(), ldo=out.stride(0),
                           accumulate=int(accumulate))
    L.check(_lib().srw_splitk_reduce(C.byref(a), _s()), "srw_splitk_reduce")
    return out


def colsum(x: torch.Tensor | None = None, planes: Planes | None = None, out: torch.Tensor | None = None, accumulate: bool = False,
           row_scale=None, rows_per_scale: int = 1):
    rows, cols = (x.shape if x is not None else (planes.rows, planes.cols))
    dev = x.device if x is not None else planes.t.device
    if out is None:
        out = torch.empty(cols, dtype=torch.float32, device=dev)
    ws = torch.empty(256 * cols, dtype=torch.float32, device=dev)
    a = L.ColsumArgs(x=L.ptr(x), ldx=x.stride(0) if x is not None else 0, planes=planes.ptr() if planes else None,
                     ldp=planes.ld if planes else 0, plane_stride=planes.plane_stride if planes else 0,
                     row_scale=L.ptr(row_scale), rows_per_scale=rows_per_scale, rows=rows, cols=cols, out=out.data_ptr(),
                     accumulate=int(accumulate), workspace=ws.data_ptr())
    L.check(_lib().srw_colsum(C.byref(a), _s()), "srw_colsum")
    return out


def layernorm_fwd(x: torch.Tensor, gamma, beta, eps: float, want_f32: bool = True, want_planes: bool = False):
    rows, cols = x.shape
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    y = torch.empty_like(x) if want_f32 else None
    yp = empty_planes(rows, cols, x.device) if want_planes else None
    a = L.LayerNormFwdArgs(x=x.data_ptr(), ldx=x.stride(0), rows=rows, cols=cols, eps=eps, gamma=gamma.data_ptr(),
                           beta=beta.data_ptr(), mean=mean.data_ptr(), rstd=rstd.data_ptr(),
                           y_planes=yp.ptr() if yp else None, ldp=yp.ld if yp else 0, plane_stride=yp.plane_stride if yp else 0,
                           y_f32=L.ptr(y), ldy=y.stride(0) if y is not None else 0)
    L.check(_lib().srw_layernorm_fwd(C.byref(a), _s()), "srw_layernorm_fwd")
    return y, yp, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dx=None, accumulate_dx=False, dgamma=None, dbeta=None, accumulate_dparams=False,
                  dx_planes: Planes | None = None, colsum_out=None, drop: L.Dropout | None = None, drop_rows_per_seq: int = 1):
    rows, cols = x.shape
    if dx is None:
        dx = torch.empty_like(x)
    if dgamma is None:
        dgamma = torch.empty(cols, dtype=torch.float32, device=x.device)
        dbeta = torch.empty_like(dgamma)
    ws = torch.empty(3 * 256 * cols, dtype=torch.float32, device=x.device)
    a = L.LayerNormBwdArgs(dy=dy.data_ptr(), lddy=dy.stride(0), x=x.data_ptr(), ldx=x.stride(0), rows=rows, cols=cols,
                           gamma=gamma.data_ptr(), mean=mean.data_ptr(), rstd=rstd.data_ptr(), dx=dx.data_ptr(), lddx=dx.stride(0),
                           accumulate_dx=int(accumulate_dx), dgamma=dgamma.data_ptr(), dbeta=dbeta.data_ptr(),
                           accumulate_dparams=int(accumulate_dparams), workspace=ws.data_ptr(),
                           dx_planes=dx_planes.ptr() if dx_planes else None, ldp=dx_planes.ld if dx_planes else 0,
                           plane_stride=dx_planes.plane_stride if dx_planes else 0, row_scale=None, rows_per_scale=1,
                           colsum_out=L.ptr(colsum_out), colsum_accumulate=0, drop=drop if drop is not None else dropout_spec(),
                           drop_rows_per_seq=drop_rows_per_seq)
    L.check(_lib().srw_layernorm_bwd(C.byref(a), _s()), "srw_layernorm_bwd")
    return dx, dgamma, dbeta


def dropout_spec(seq_key=None, seq_row=None, site: int = 0, p: float = 0.0) -> L.Dropout:
    """srw_dropout: seq_key uint32 (stored as int32/int64 tensors are NOT accepted: pass torch.int32 views of uint32 bits) [S],
    seq_row int32 [S]; None / p = 0 -> off."""
    if seq_key is None or p == 0.0:
        return L.Dropout(seq_key=None, seq_row=None, site=0, p=0.0)
    assert seq_key.dtype == torch.int32 and seq_row.dtype == torch.int32 and seq_key.is_contiguous() and seq_row.is_contiguous()
    return L.Dropout(seq_key=seq_key.data_ptr(), seq_row=seq_row.data_ptr(), site=int(site), p=float(p))


def attn_mask_prepare(attention_mask, B: int, Lq: int, device="cuda"):
    """attention_mask int64 [B, L] or None -> (key_bias fp32 [B, ld], kv_len int32 [B])."""
    ld = (Lq + 63) // 64 * 64
    bias = torch.empty(B, ld, dtype=torch.float32, device=device)
    kv = torch.empty(B, dtype=torch.int32, device=device)
    L.check(_lib().srw_attn_mask_prepare(L.ptr(attention_mask), B, Lq, bias.data_ptr(), ld, kv.data_ptr(), _s()), "srw_attn_mask_prepare")
    return bias, kv


def attn_fwd(qkv: Planes, B: int, N: int, H: int, head_dim: int = 64, key_bias=None, kv_len=None, drop: L.Dropout | None = None):
    """qkv planes [B*N, 3*H*head_dim] -> (o planes [B*N, H*head_dim], lse [B,H,N])."""
    dev = qkv.t.device
    D = H * head_dim
    o = empty_planes(B * N, D, dev)
    lse = torch.empty(B, H, N, dtype=torch.float32, device=dev)
    a = L.AttnFwdArgs(B=B, N=N, H=H, head_dim=head_dim, scale=head_dim ** -0.5, qkv=qkv.ptr(), ld_qkv=qkv.ld,
                      qkv_plane_stride=qkv.plane_stride, o=o.ptr(), ld_o=o.ld, o_plane_stride=o.plane_stride, lse=lse.data_ptr(),
                      key_bias=L.ptr(key_bias), ld_bias=key_bias.stride(0) if key_bias is not None else 0, kv_len=L.ptr(kv_len),
                      drop=drop if drop is not None else dropout_spec())
    L.check(_lib().srw_attn_fwd(C.byref(a), _s()), "srw_attn_fwd")
    return o, lse


def attn_bwd(qkv: Planes, o: Planes, d_o: Planes, lse: torch.Tensor, B: int, N: int, H: int, head_dim: int = 64, key_bias=None, kv_len=None,
             drop: L.Dropout | None = None):
    dev = qkv.t.device
    D = H * head_dim
    dqkv = empty_planes(B * N, 3 * D, dev)
    delta = torch.empty(B, H, N, dtype=torch.float32, device=dev)
    a = L.AttnBwdArgs(B=B, N=N, H=H, head_dim=head_dim, scale=head_dim ** -0.5, qkv=qkv.ptr(), ld_qkv=qkv.ld,
                      qkv_plane_stride=qkv.plane_stride, o=o.ptr(), ld_o=o.ld, o_plane_stride=o.plane_stride,
                      d_o=d_o.ptr(), ld_do=d_o.ld, do_plane_stride=d_o.plane_stride, lse=lse.data_ptr(), delta=delta.data_ptr(),
                      dqkv=dqkv.ptr(), ld_dqkv=dqkv.ld, dqkv_plane_stride=dqkv.plane_stride,
                      key_bias=L.ptr(key_bias), ld_bias=key_bias.stride(0) if key_bias is not None else 0, kv_len=L.ptr(kv_len),
                      drop=drop if drop is not None else dropout_spec())
    L.check(_lib().srw_attn_bwd(C.byref(a), _s()), "srw_attn_bwd")
    return dqkv
