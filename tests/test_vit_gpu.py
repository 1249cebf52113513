"""-m gpu: the native ViT engine (srw_vit_forward/backward through the nn.Module boundary) against the oracle's
vit_forward + torch autograd on CPU fp32, same deterministic weights and images.
Tolerances (BASELINE.json north_star): logits/feat within 1e-3 abs; gradients within 1e-3 relative to the tensor's
largest entry."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(depth, num_classes=100, head_gain=4.0, seed=0):
    from oracle import ssl_oracle as O
    from semireward_b200 import detgen
    from semireward_b200.nets import vit_small_patch2_32
    vc = O.ViTConfig(depth=depth, num_classes=num_classes)
    p = {n: torch.from_numpy(detgen.fill_param(n, s, seed)) for n, s in vc.param_shapes()}
    p["head.weight"] = p["head.weight"] * head_gain
    model = vit_small_patch2_32(num_classes=num_classes, depth=depth, drop_path_rate=0.0)
    model.load_state_dict(p)
    model = model.cuda().train()
    return O, vc, p, model


@pytest.mark.parametrize("depth,use_drop", [(2, False), (2, True), (12, False)])
def test_vit_forward_backward_vs_oracle(depth, use_drop):
    from semireward_b200 import detgen
    O, vc, p, model = _setup(depth)
    batch = detgen.ssl_batch(8, 1, 100, 50000, seed=1, step=0)
    # engine order: rows that carry gradient first (labelled, strong), weak rows last
    x = torch.from_numpy(np.concatenate([batch["x_lb"], batch["x_ulb_s"], batch["x_ulb_w"]]))
    B, Bg = x.shape[0], 16
    drop = None
    if use_drop:
        g = torch.Generator().manual_seed(5)
        drop = (torch.bernoulli(torch.full((depth, 2, B), 0.7), generator=g) / 0.7)
    out = model(x.cuda(), grad_batch=Bg, drop_scale=drop.cuda() if drop is not None else None)
    logits, feat = out["logits"], out["feat"]
    # oracle
    po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lo, fo = O.vit_forward(po, x, vc, drop)
    err_l = (logits.cpu() - lo.detach()).abs().max().item()
    err_f = (feat.cpu() - fo.detach()).abs().max().item()
    print(f"depth {depth} drop {use_drop}: logits err {err_l:.3e} (max |logit| {lo.abs().max().item():.2f}), feat err {err_f:.3e}")
    assert err_l < 1e-3 and err_f < 1e-3
    # backward: random cotangents on the gradient-carrying rows only
    g = torch.Generator().manual_seed(7)
    cl = torch.randn(B, vc.num_classes, generator=g)
    cf = torch.randn(B, vc.embed_dim, generator=g) * 0.1
    cl[Bg:] = 0
    cf[Bg:] = 0
    loss = (logits * cl.cuda()).sum() + (feat * cf.cuda()).sum()
    loss.backward()
    ((lo * cl).sum() + (fo * cf).sum()).backward()
    worst = 0.0
    for name, prm in model.named_parameters():
        ref = po[name].grad
        got = prm.grad.cpu()
        scale = ref.abs().max().item()
        err = (got - ref).abs().max().item()
        rel = err / max(scale, 1e-12)
        worst = max(worst, rel)
        assert rel < 1e-3, f"{name}: grad err {err:.3e} vs max |grad| {scale:.3e}"
    print(f"worst relative gradient error {worst:.3e}")


def test_vit_simt_twin_matches_tcgen05():
    """The SIMT verification GEMM (fp32 FMA over the same planes) and the tcgen05 path agree to fp32-accumulate level."""
    from semireward_b200 import _lib as L, detgen
    O, vc, p, model = _setup(2)
    x = torch.from_numpy(detgen.normal("x", (4, 3, 32, 32), 3)).cuda()
    with torch.no_grad():
        a = model(x)["logits"]
        model.gemm_impl = L.GEMM_SIMT
        b = model(x)["logits"]
    assert (a - b).abs().max().item() < 1e-3


@pytest.mark.parametrize("use_drop", [False, True])
def test_native_pass_graph_replay_matches_autograd_path(use_drop):
    """forward_native/backward_native (persistent buffers -> CUDA-graph replay from the 3rd identical call on) give
    bit-identical logits and gradients to the autograd route, call after call, and follow changing inputs."""
    from semireward_b200 import _lib as L, detgen
    O, vc, p, model = _setup(2)
    lib = L.load()
    B, Bg = 12, 8
    for rep in range(5):   # rep 0 eager, rep 1 captures, rep >= 2 replays
        x = torch.from_numpy(detgen.normal("x", (B, 3, 32, 32), 10 + rep)).cuda()
        drop = None
        if use_drop:
            g = torch.Generator().manual_seed(rep)
            drop = (torch.bernoulli(torch.full((2, 2, B), 0.7), generator=g) / 0.7).cuda()
        cl = torch.from_numpy(detgen.normal("cl", (Bg, 100), 20 + rep)).cuda()
        model.zero_grad()
        out = model(x, grad_batch=Bg, drop_scale=drop)
        (out["logits"][:Bg] * cl).sum().backward()
        ref_grads = [q.grad.clone() for q in model._ordered_params()]
        ds = None
        if drop is not None:
            ds = model._buf("test_drop", drop.shape, drop.device)
            ds.copy_(drop)
        n0 = lib.srw_kernel_launches()
        lg, ft, h = model.forward_native(x, grad_batch=Bg, drop_scale=ds)
        flat, views = model.backward_native(h, cl)
        assert lib.srw_kernel_launches() - n0 > 50   # replays are counted like launches
        torch.cuda.synchronize()
        assert torch.equal(lg, out["logits"].detach()) and torch.equal(ft, out["feat"].detach()), f"rep {rep}"
        for a, b in zip(views, ref_grads):
            assert torch.equal(a, b), f"rep {rep}: gradient differs between graph replay and eager launches"


def test_scale_inplace():
    from semireward_b200 import _lib as L
    x = torch.arange(1003, dtype=torch.float32, device="cuda")
    one, half = torch.ones((), device="cuda"), torch.full((), 0.5, device="cuda")
    lib = L.load()
    L.check(lib.srw_scale_inplace(x.data_ptr(), x.numel(), one.data_ptr(), L.stream_ptr()))
    assert torch.equal(x, torch.arange(1003, dtype=torch.float32, device="cuda"))
    L.check(lib.srw_scale_inplace(x.data_ptr(), x.numel(), half.data_ptr(), L.stream_ptr()))
    assert torch.equal(x, torch.arange(1003, dtype=torch.float32, device="cuda") * 0.5)


@pytest.mark.parametrize("builder,kw,vc_kw", [
    ("vit_tiny_patch2_32", dict(), dict(img_size=32, patch_size=2, embed_dim=192, num_heads=3)),
    ("vit_base_patch16_96", dict(), dict(img_size=96, patch_size=16, embed_dim=768, num_heads=12)),
    ("vit_small_patch16_224", dict(), dict(img_size=224, patch_size=16, embed_dim=384, num_heads=6)),
])
def test_other_vit_builders_vs_oracle(builder, kw, vc_kw):
    """The remaining ViT builders of semilearn/nets/vit/vit.py:323-408 through the same engine:
    D = 192 / 768, N = 37 / 197 / 257 tokens, patch-embed K = 12 / 768."""
    from oracle import ssl_oracle as O
    from semireward_b200 import detgen, nets
    depth, C = 2, 10
    vc = O.ViTConfig(depth=depth, num_classes=C, **vc_kw)
    p = {n: torch.from_numpy(detgen.fill_param(n, s, 0)) for n, s in vc.param_shapes()}
    p["head.weight"] = p["head.weight"] * 4.0
    model = getattr(nets, builder)(num_classes=C, depth=depth, drop_path_rate=0.0, **kw)
    model.load_state_dict(p)
    model = model.cuda().train()
    B, Bg = 6, 4
    x = torch.from_numpy(detgen.normal("x_builders", (B, 3, vc.img_size, vc.img_size), 5))
    out = model(x.cuda(), grad_batch=Bg)
    po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lo, fo = O.vit_forward(po, x, vc, None)
    assert (out["logits"].cpu() - lo.detach()).abs().max().item() < 1e-3
    assert (out["feat"].cpu() - fo.detach()).abs().max().item() < 1e-3
    cl = torch.from_numpy(detgen.normal("cl_builders", (B, C), 6))
    cl[Bg:] = 0
    (out["logits"] * cl.cuda()).sum().backward()
    (lo * cl).sum().backward()
    for name, prm in model.named_parameters():
        ref = po[name].grad
        rel = (prm.grad.cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)
        assert rel < 1e-3, f"{builder} {name}: relative gradient error {rel:.3e}"


def test_split_backward_equals_full_backward():
    """srw_vit_backward over two descending block ranges (the data-parallel overlap route) == one full call, bit for bit;
    the all-reduce hand-off is stubbed so that this runs in one process."""
    from semireward_b200 import detgen
    O, vc, p, model = _setup(4)
    B, Bg = 9, 6
    x = torch.from_numpy(detgen.normal("x_split", (B, 3, 32, 32), 31)).cuda()
    cl = torch.from_numpy(detgen.normal("cl_split", (Bg, 100), 32)).cuda()

    class _Done:
        def wait(self):
            return True
    calls = []
    results = []
    for split in (0, 2, 3, 4):   # number of block ranges (depth 4)
        model.dp_overlap_split = split
        model._dp_group = object() if split else None
        model._allreduce_async = lambda t, g: (calls.append(t.numel()) or (_Done(), None))
        for rep in range(3):   # eager, captured, replayed
            lg, ft, h = model.forward_native(x, grad_batch=Bg)
            flat, views = model.backward_native(h, cl)
            if split:
                model.allreduce_grads_()
        torch.cuda.synchronize()
        results.append(flat.clone())
    model._dp_group = None
    assert all(torch.equal(results[0], r) for r in results[1:])
    assert len(calls) == 3 * (2 + 3 + 4) and sum(calls[:2]) == results[0].numel()


@pytest.mark.parametrize("harsh", [False, True])
def test_vit_pretrained_scale_statistics(harsh):
    """the 1e-3 logit gate had only seen hash-filled weights (|logit| <= 24).  Pre-trained ViTs have heavier statistics:
    LayerNorm gains far from 1, a few outlier channels in the residual stream, wider MLP weights, logits of several tens.
    harsh = False: gains in [0.5, 2] with x2 outlier channels, weights x1.3, |logit| ~ 40: the gates stay 1e-3 ABSOLUTE on logits and
    features, 1e-3 relative on gradients.
    harsh = True: gains to 3 with x4 outliers, weights x1.6, position embeddings x10: attention scores of several hundred, i.e. saturated
    softmaxes.  The split-bf16 products carry 16 mantissa bits per operand: a score S enters exp() with an absolute error
    of ~1.5e-5 |S|, which a saturated softmax amplifies; this case documents the limit (relative logit error printed, held to 2e-3 =
    the accuracy of a single bf16x3 product chain through 12 saturated blocks) instead of pretending the fp32 gate holds there."""
    from oracle import ssl_oracle as O
    from semireward_b200 import detgen
    from semireward_b200.nets import vit_small_patch2_32
    vc = O.ViTConfig(depth=12, num_classes=100)
    g = torch.Generator().manual_seed(17)
    gain_hi, outlier, wscale, pos_scale = (2.5, 4.0, 1.6, 10.0) if harsh else (1.5, 2.0, 1.3, 3.0)
    p = {}
    for n, s in vc.param_shapes():
        t = torch.from_numpy(detgen.fill_param(n, s, 0))
        if n.endswith(("norm1.weight", "norm2.weight")) or n == "norm.weight":
            t = t * (0.5 + gain_hi * torch.rand(t.shape, generator=g))
            t[torch.randint(0, t.numel(), (3,), generator=g)] *= outlier      # a few outlier channels
        elif n.endswith(("fc1.weight", "fc2.weight", "qkv.weight")):
            t = t * wscale
        elif n == "pos_embed":
            t = t * pos_scale
        elif n == "head.weight":
            t = t * 3.5
        p[n] = t
    model = vit_small_patch2_32(num_classes=100, depth=12, drop_path_rate=0.0)
    model.load_state_dict(p)
    model = model.cuda().train()
    batch = detgen.ssl_batch(8, 1, 100, 50000, seed=1, step=0)
    x = torch.from_numpy(np.concatenate([batch["x_lb"], batch["x_ulb_s"], batch["x_ulb_w"]]))
    B, Bg = x.shape[0], 16
    out = model(x.cuda(), grad_batch=Bg)
    po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    lo, fo = O.vit_forward(po, x, vc, None)
    lmax = lo.abs().max().item()
    err_l = (out["logits"].cpu() - lo.detach()).abs().max().item()
    err_f = (out["feat"].cpu() - fo.detach()).abs().max().item()
    print(f"pretrained-scale statistics (harsh {harsh}): max |logit| {lmax:.1f}, max |feat| {fo.abs().max().item():.1f}: logits err {err_l:.3e} "
          f"({err_l / lmax:.1e} relative), feat err {err_f:.3e}")
    if harsh:
        assert err_l < 2e-3 * lmax and err_f < 2e-3 * fo.abs().max().item()
        return
    assert lmax > 30.0, "the fixture is meant to produce large logits"
    assert err_l < 1e-3 and err_f < 1e-3
    gen = torch.Generator().manual_seed(7)
    cl = torch.randn(B, 100, generator=gen)
    cl[Bg:] = 0
    (out["logits"] * cl.cuda()).sum().backward()
    (lo * cl).sum().backward()
    worst, wn = 0.0, ""
    for name, prm in model.named_parameters():
        ref = po[name].grad
        rel = (prm.grad.cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)
        if rel > worst:
            worst, wn = rel, name
    print(f"   worst relative gradient error {worst:.3e} ({wn})")
    assert worst < 1e-3, (worst, wn)
