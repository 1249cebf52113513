"""-m gpu: single-kernel parity of the C-ABI kernels against a float64 torch restatement of the same op.
Tolerances are stated per test; the bf16x3 tensor-core GEMM is held to fp32-sgemm accuracy (a few 1e-6 relative)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from semireward_b200 import _lib as L, ops as O
    lib = L.load()
    import ctypes as C
    ma, mi, sm = C.c_int(), C.c_int(), C.c_int()
    L.check(lib.srw_device_check(C.byref(ma), C.byref(mi), C.byref(sm)), "srw_device_check")
    return O


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def test_split_planes_roundtrip(ops):
    x = _rand(300, 392, seed=1)
    p, pt = ops.split_planes(x, transposed=True)
    torch.cuda.synchronize()
    err = (p.to_f32() - x).abs().max().item()
    assert err <= 2 ** -16 * x.abs().max().item()
    assert torch.equal(pt.to_f32(), p.to_f32().t())


@pytest.mark.parametrize("impl", [1, 2, 0])  # SIMT twin, 1-CTA tcgen05, default (2-CTA pairs where they fit)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (6168, 1152, 384), (257, 100, 1536), (384, 384, 6168), (1000, 64, 200), (300, 1536, 384)])
def test_gemm_f32(ops, impl, a_mn, b_mn, M, N, K):
    from semireward_b200 import _lib as L
    A = _rand(M, K, seed=2)
    B = _rand(N, K, seed=3)
    if a_mn and M % 8:
        pytest.skip("MN-major operand needs ld % 8 == 0")
    if b_mn and N % 8:
        pytest.skip("MN-major operand needs ld % 8 == 0")
    pa = ops.split_planes(A.t().contiguous()) if a_mn else ops.split_planes(A)
    pb = ops.split_planes(B.t().contiguous()) if b_mn else ops.split_planes(B)
    bias = _rand(N, seed=4)
    if K > 2048 and impl != 1:
        # one fp32 TMEM accumulator per tile: long reductions go through split-K (as the engine's wgrad GEMMs do)
        with pytest.raises(L.SrwError, match="K per CTA"):
            ops.gemm(pa, pb, M, N, K, a_mn=a_mn, b_mn=b_mn, epilogue=L.EPI_F32, bias=bias, impl=impl)
        ws = ops.gemm(pa, pb, M, N, K, a_mn=a_mn, b_mn=b_mn, epilogue=L.EPI_SPLITK, split_k=8, impl=impl)
        out = bias.repeat(M, 1).contiguous()
        ops.splitk_reduce(ws, out, accumulate=True)
    else:
        out, _ = ops.gemm(pa, pb, M, N, K, a_mn=a_mn, b_mn=b_mn, epilogue=L.EPI_F32, bias=bias, impl=impl)
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().t() + bias.double())
    err = (out.double() - ref).abs().max().item()
    tol = 6e-5 * (K ** 0.5)  # split-bf16 operands carry ~2^-17 relative error each; N(0,1) data
    assert err < tol, f"max abs err {err} (tol {tol})"


def test_gemm_epilogues(ops):
    from semireward_b200 import _lib as L
    M, N, K = 771, 384, 384
    A, B, bias = _rand(M, K, seed=5), _rand(N, K, seed=6, scale=0.05), _rand(N, seed=7)
    pa, pb = ops.split_planes(A), ops.split_planes(B)
    z_ref = A.double() @ B.double().t() + bias.double()
    # PLANES
    _, op = ops.gemm(pa, pb, M, N, K, epilogue=L.EPI_PLANES, bias=bias)
    assert (op.to_f32().double() - z_ref).abs().max().item() < 2e-4
    # GELU: out_f32 = z, planes = gelu(z)
    z, gp = ops.gemm(pa, pb, M, N, K, epilogue=L.EPI_GELU, bias=bias)
    assert (z.double() - z_ref).abs().max().item() < 2e-4
    assert (gp.to_f32().double() - torch.nn.functional.gelu(z_ref)).abs().max().item() < 2e-4
    # RESID with per-image row scale (257 rows per image)
    resid = _rand(M, N, seed=8)
    rs = torch.tensor([1.0, 0.0, 1.25], device="cuda")
    o, _ = ops.gemm(pa, pb, M, N, K, epilogue=L.EPI_RESID, bias=bias, resid=resid, row_scale=rs, rows_per_scale=257)
    ref = resid.double() + rs.double().repeat_interleave(257)[:, None] * z_ref
    assert (o.double() - ref).abs().max().item() < 3e-4
    # DGELU: planes = acc * gelu'(aux)
    aux = _rand(M, N, seed=9)
    _, dp = ops.gemm(pa, pb, M, N, K, epilogue=L.EPI_DGELU, aux=aux)
    ax = aux.double().requires_grad_(True)
    torch.nn.functional.gelu(ax).sum().backward()
    ref = (A.double() @ B.double().t()) * ax.grad
    assert (dp.to_f32().double() - ref).abs().max().item() < 3e-4
    # split-K
    ws = ops.gemm(pa, pb, M, N, K, epilogue=L.EPI_SPLITK, split_k=3)
    out = torch.zeros(M, N, device="cuda")
    ops.splitk_reduce(ws, out)
    assert (out.double() - A.double() @ B.double().t()).abs().max().item() < 3e-4


def test_colsum(ops):
    x = _rand(6168, 384, seed=10)
    out = ops.colsum(x=x)
    assert (out.double() - x.double().sum(0)).abs().max().item() < 2e-3
    p = ops.split_planes(x)
    out2 = ops.colsum(planes=p)
    assert (out2.double() - x.double().sum(0)).abs().max().item() < 2e-3


def test_layernorm_fwd_bwd(ops):
    rows, cols = 1031, 384
    x = _rand(rows, cols, seed=11, scale=2.0) + 0.3
    g, b = _rand(cols, seed=12) * 0.1 + 1.0, _rand(cols, seed=13) * 0.1
    y, yp, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-6, want_f32=True, want_planes=True)
    xd = x.double().requires_grad_(True)
    gd, bd = g.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xd, (cols,), gd, bd, 1e-6)
    assert (y.double() - ref).abs().max().item() < 5e-6
    assert (yp.to_f32().double() - ref).abs().max().item() < 6e-5
    dy = _rand(rows, cols, seed=14)
    ref.backward(dy.double())
    dx, dg, db = ops.layernorm_bwd(dy, x, g, mean, rstd)
    assert (dx.double() - xd.grad).abs().max().item() < 2e-5
    assert (dg.double() - gd.grad).abs().max().item() < 1e-3
    assert (db.double() - bd.grad).abs().max().item() < 1e-3


def _attn_ref(qkv, B, N, H):
    D = H * 64
    x = qkv.double().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    q, k, v = x[0], x[1], x[2]
    s = (q @ k.transpose(-2, -1)) * 64 ** -0.5
    a = s.softmax(-1)
    o = (a @ v).transpose(1, 2).reshape(B * N, D)
    return o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,N,H", [(3, 257, 6), (2, 197, 2), (1, 64, 1), (2, 272, 1), (1, 16, 1), (2, 129, 2), (1, 256, 1), (2, 65, 1), (1, 144, 2)])
def test_attention_fwd_bwd(ops, B, N, H):
    D = H * 64
    qkv = _rand(B * N, 3 * D, seed=20, scale=1.5)
    pq = ops.split_planes(qkv)
    o, lse = ops.attn_fwd(pq, B, N, H)
    torch.cuda.synchronize()
    qd = pq.to_f32().double().requires_grad_(True)   # the kernel sees the split-plane values
    o_ref, lse_ref = _attn_ref(qd, B, N, H)
    err_o = (o.to_f32().double() - o_ref).abs().max().item()
    err_l = (lse.double() - lse_ref).abs().max().item()
    assert err_o < 5e-5, f"o err {err_o}"
    assert err_l < 2e-5, f"lse err {err_l}"
    d_o = _rand(B * N, D, seed=21)
    pdo = ops.split_planes(d_o)
    dqkv = ops.attn_bwd(pq, o, pdo, lse, B, N, H)
    torch.cuda.synchronize()
    o_ref.backward(pdo.to_f32().double())
    err_g = (dqkv.to_f32().double() - qd.grad).abs().max().item()
    scale_g = qd.grad.abs().max().item()
    assert err_g < 1e-4 * max(1.0, scale_g), f"dqkv err {err_g} (max |grad| {scale_g})"


def test_grad_fold_matches_separate_reductions(ops):
    """srw_grad_fold (one launch per block) against float64 sums and, bit for bit, against srw_splitk_reduce."""
    import ctypes as C
    from semireward_b200 import _lib as L
    lib = L.load()
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    a = L.GradFoldArgs()
    keep, checks = [], []
    shapes = [(384, 1536, 5, 1536, 0), (1152, 384, 8, 384, 1), (384, 384, 3, 392, 1)]   # (M, N, split, ldo, accumulate)
    for i, (M, N, split, ldo, acc) in enumerate(shapes):
        ws = _rand(split, M, N, seed=30 + i)
        out = _rand(M, ldo, seed=40 + i)
        ref = (ws.double().sum(0) + (out[:, :N].double() if acc else 0))
        out_sep = out.clone()
        r = L.SplitKReduceArgs(workspace=ws.data_ptr(), split_k=split, M=M, N=N, out=out_sep.data_ptr(), ldo=ldo, accumulate=acc)
        L.check(lib.srw_splitk_reduce(C.byref(r), s), "srw_splitk_reduce")
        a.splitk[i] = L.SplitKReduceArgs(workspace=ws.data_ptr(), split_k=split, M=M, N=N, out=out.data_ptr(), ldo=ldo, accumulate=acc)
        keep += [ws, out]
        checks.append((out, ref, out_sep, N))
    a.n_splitk = len(shapes)
    csum = [(64, 1536, 1536, 0), (129, 384, 3 * 384, 1), (7, 100, 100, 0)]   # (nparts, cols, stride_p, accumulate)
    cchecks = []
    for i, (nparts, cols, stride, acc) in enumerate(csum):
        part = _rand(nparts, stride, seed=50 + i)
        out = _rand(cols, seed=60 + i)
        ref = part[:, :cols].double().sum(0) + (out.double() if acc else 0)
        a.colsum[i] = L.FoldColsum(partial=part.data_ptr(), nparts=nparts, stride_p=stride, cols=cols, out=out.data_ptr(), accumulate=acc)
        keep += [part, out]
        cchecks.append((out, ref))
    a.n_colsum = len(csum)
    L.check(lib.srw_grad_fold(C.byref(a), s), "srw_grad_fold")
    torch.cuda.synchronize()
    for out, ref, out_sep, N in checks:
        assert torch.equal(out, out_sep), "fold differs from srw_splitk_reduce"
        assert (out[:, :N].double() - ref).abs().max().item() < 1e-5
    for out, ref in cchecks:
        assert (out.double() - ref).abs().max().item() < 1e-4
    assert lib.srw_colsum_nparts(4112) == 64 and lib.srw_layernorm_bwd_nparts(4112) == 256 and lib.srw_layernorm_bwd_nparts(100) == 13
