"""-m gpu: the kernel features the text / audio encoders add (HF BertModel / HubertModel behind bert.py:34, hubert.py:45), each
against a float64 torch restatement with the SAME counter-based dropout bits (oracle/bert_oracle.counter_keep_mask):
key-streaming attention up to 512 tokens with a key-padding mask and dropout on the probabilities (forward and backward),
dropout in the residual GEMM epilogue, dropout in the LayerNorm-backward hand-over."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from semireward_b200 import ops as O
    return O


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def _i32(vals):
    return torch.from_numpy(np.asarray(vals, dtype=np.uint32).view(np.int32).copy()).cuda()


def _keep(numel, keep, key, site):
    from oracle import bert_oracle as BO
    return BO.counter_keep_mask(numel, keep, int(key), site).cuda()


@pytest.mark.parametrize("B,N,H,masked,p", [(2, 512, 2, True, 0.1), (3, 512, 1, True, 0.0), (2, 199, 2, False, 0.1), (1, 384, 1, False, 0.0),
                                            (2, 130, 1, True, 0.25), (1, 512, 12, False, 0.0), (2, 64, 1, True, 0.1)])
def test_streaming_attention_mask_dropout_fwd_bwd(ops, B, N, H, masked, p):
    D = H * 64
    qkv = _rand(B * N, 3 * D, seed=20, scale=1.5)
    pq = ops.split_planes(qkv)
    am = None
    if masked:   # trailing padding of different lengths, plus one hole inside a sequence (general masks are allowed)
        am = torch.ones(B, N, dtype=torch.long)
        lens = [N - 7 - 61 * i for i in range(B)]
        for i, ln in enumerate(lens):
            am[i, max(ln, 17):] = 0
        am[0, 5] = 0
        am = am.cuda()
    bias, kv = (ops.attn_mask_prepare(am, B, N) if masked else (None, None))
    keys, rows = [0x9E3779B1 * (i + 3) & 0xFFFFFFFF for i in range(B)], [(5 * i + 2) % 7 for i in range(B)]
    tk, trw = _i32(keys), torch.tensor(rows, dtype=torch.int32).cuda()   # the spec only borrows their device pointers: keep them alive
    drop = ops.dropout_spec(tk, trw, site=4, p=p) if p else None
    o, lse = ops.attn_fwd(pq, B, N, H, key_bias=bias, kv_len=kv, drop=drop)
    torch.cuda.synchronize()
    if masked:
        assert kv.tolist() == [int(am[i].nonzero().max()) + 1 for i in range(B)]
    qd = pq.to_f32().double().requires_grad_(True)
    x = qd.reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (x[0] @ x[1].transpose(-2, -1)) * 64 ** -0.5
    if masked:
        s = s.masked_fill(am[:, None, None, :] == 0, float("-inf"))
    a = s.softmax(-1)
    if p:
        keepm = torch.stack([_keep_rows(rows[i], H, N, 1 - p, keys[i], 4) for i in range(B)])
        a = a * keepm.double() / (1 - p)
    o_ref = (a @ x[2]).transpose(1, 2).reshape(B * N, D)
    lse_ref = torch.logsumexp(s, -1)
    err_o = (o.to_f32().double() - o_ref).abs().max().item()
    err_l = (lse.double() - lse_ref).abs().max().item()
    assert err_o < 8e-5, f"o err {err_o}"
    assert err_l < 3e-5, f"lse err {err_l}"
    d_o = _rand(B * N, D, seed=21)
    pdo = ops.split_planes(d_o)
    dqkv = ops.attn_bwd(pq, o, pdo, lse, B, N, H, key_bias=bias, kv_len=kv, drop=drop)
    torch.cuda.synchronize()
    o_ref.backward(pdo.to_f32().double())
    err_g = (dqkv.to_f32().double() - qd.grad).abs().max().item()
    scale_g = qd.grad.abs().max().item()
    assert err_g < 1e-4 * max(1.0, scale_g), f"dqkv err {err_g} (max |grad| {scale_g})"


def _keep_rows(seq_row, H, N, keep, key, site):
    """keep bits of one sequence's attention probabilities [H, N, N]: element ((seq_row * H + h) * N + q) * N + k of the call's tensor"""
    from oracle import bert_oracle as BO
    return BO.counter_keep_mask(H * N * N, keep, int(key), site, offset=seq_row * H * N * N).view(H, N, N).cuda()


def test_resid_epilogue_dropout(ops):
    from semireward_b200 import _lib as L
    Lq, S, N, K = 96, 5, 768, 384           # 5 sequences of 96 tokens: M = 480 (interior and edge tiles)
    M = Lq * S
    A, B, bias, resid = _rand(M, K, seed=5), _rand(N, K, seed=6, scale=0.05), _rand(N, seed=7), _rand(M, N, seed=8)
    pa, pb = ops.split_planes(A), ops.split_planes(B)
    keys, rows = [11, 11, 11, 977, 977], [0, 1, 2, 0, 1]       # two "calls": rows 0..2 of stream 11, rows 0..1 of stream 977
    tk, trw = _i32(keys), torch.tensor(rows, dtype=torch.int32).cuda()
    drop = ops.dropout_spec(tk, trw, site=2, p=0.1)
    out, _ = ops.gemm(pa, pb, M, N, K, epilogue=L.EPI_RESID, bias=bias, resid=resid, drop=drop, drop_rows_per_seq=Lq)
    torch.cuda.synchronize()
    z = A.double() @ B.double().t() + bias.double()
    keepm = torch.cat([_keep(3 * Lq * N, 0.9, 11, 2), _keep(2 * Lq * N, 0.9, 977, 2)]).view(M, N)
    ref = resid.double() + z * keepm.double() / 0.9
    assert (out.double() - ref).abs().max().item() < 3e-4
    assert abs(float(keepm.float().mean()) - 0.9) < 5e-3


def test_layernorm_bwd_dropout_handover(ops):
    Lq, S, cols = 40, 3, 768
    rows = Lq * S
    x = _rand(rows, cols, seed=11, scale=2.0) + 0.3
    g, b = _rand(cols, seed=12) * 0.1 + 1.0, _rand(cols, seed=13) * 0.1
    y, _, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12, want_f32=True)
    dy = _rand(rows, cols, seed=14)
    keys, srows = [123456789, 123456789, 42], [4, 5, 0]
    tk, trw = _i32(keys), torch.tensor(srows, dtype=torch.int32).cuda()
    drop = ops.dropout_spec(tk, trw, site=9, p=0.1)
    dxp = ops.empty_planes(rows, cols)
    cs = torch.empty(cols, device="cuda")
    dx, dg, db = ops.layernorm_bwd(dy, x, g, mean, rstd, dx_planes=dxp, colsum_out=cs, drop=drop, drop_rows_per_seq=Lq)
    torch.cuda.synchronize()
    xd = x.double().requires_grad_(True)
    torch.nn.functional.layer_norm(xd, (cols,), g.double(), b.double(), 1e-12).backward(dy.double())
    assert (dx.double() - xd.grad).abs().max().item() < 2e-5
    from oracle import bert_oracle as BO
    keepm = torch.cat([BO.counter_keep_mask(2 * Lq * cols, 0.9, 123456789, 9, offset=4 * Lq * cols), BO.counter_keep_mask(Lq * cols, 0.9, 42, 9)]).view(rows, cols).cuda()
    want = xd.grad * keepm.double() / 0.9
    assert (dxp.to_f32().double() - want).abs().max().item() < 3e-5
    assert (cs.double() - want.sum(0)).abs().max().item() < 2e-3
