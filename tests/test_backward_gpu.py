"""Parity of the backward kernels (gradient router, BN backward, dgrad, wgrad) against torch autograd on the device."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("N,H,W,Cc,kinds", [
    (2, 16, 16, 32, ["same"]),
    (2, 16, 16, 32, ["same", "same", "pool", "same"]),
    (3, 8, 12, 64, ["same", "up"]),
    (2, 6, 6, 256, ["up"]),
    (2, 12, 20, 128, ["pool", "same", "up"]),
    (2, 16, 16, 32, ["same", "head"]),
    (1, 6, 6, 512, ["up"]),
])
def test_grad_router(N, H, W, Cc, kinds):
    import _lib as L
    import ops
    gen = torch.Generator(device=DEV).manual_seed(5 + N + Cc)
    y = (torch.randn(N, H, W, Cc, generator=gen, device=DEV)).to(torch.bfloat16)
    # quantise so that 2x2-window ties among positive activations occur
    y = (y.float() * 4).round().div(4).to(torch.bfloat16)
    scale = torch.randn(Cc, generator=gen, device=DEV)
    shift = torch.randn(Cc, generator=gen, device=DEV) * 0.3
    mean = torch.randn(Cc, generator=gen, device=DEV) * 0.1
    invstd = torch.rand(Cc, generator=gen, device=DEV) + 0.5
    z = (y.float().permute(0, 3, 1, 2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).requires_grad_(True)
    a = torch.relu(z)
    cons, loss = [], 0.0
    dlogits = wh = None
    for kind in kinds:
        if kind == "head":
            K = 3
            wh = torch.randn(K, Cc, generator=gen, device=DEV)
            dlogits = torch.randn(N, K, H, W, generator=gen, device=DEV)
            loss = loss + (F.conv2d(a, wh.view(K, Cc, 1, 1)) * dlogits).sum()
            continue
        hs, ws = {"same": (H, W), "pool": (H // 2, W // 2), "up": (2 * H, 2 * W)}[kind]
        pitch, coff = Cc + 64, 32
        dx = torch.randn(N, hs, ws, pitch, generator=gen, device=DEV).to(torch.bfloat16)
        routed = {"same": a, "pool": F.max_pool2d(a, 2, 2), "up": F.interpolate(a, scale_factor=2, mode="bilinear", align_corners=True)}[kind]
        loss = loss + (routed * dx[..., coff:coff + Cc].float().permute(0, 3, 1, 2)).sum()
        cons.append(ops.Consumer(dx, pitch, coff, {"same": L.SRC_SAME, "pool": L.SRC_POOL, "up": L.SRC_UP}[kind]))
    loss.backward()
    gz_ref = z.grad  # NCHW fp32
    gz, part = ops.grad_router(N, H, W, Cc, y, scale, shift, mean, invstd, cons, dlogits=dlogits, wh=wh)
    torch.cuda.synchronize()
    got = gz.float().permute(0, 3, 1, 2)
    err = (got - gz_ref).abs()
    assert bool((err <= 0.008 * gz_ref.abs() + 1e-3 * gz_ref.abs().max()).all()), err.max()
    yhat = (y.float().permute(0, 3, 1, 2) - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
    s1 = gz_ref.double().sum((0, 2, 3))
    s2 = (gz_ref.double() * yhat.double()).sum((0, 2, 3))
    g1 = part[:, 0].double().sum(0)
    g2 = part[:, 1].double().sum(0)
    tol = 1e-3 * float(gz_ref.abs().sum((0, 2, 3)).max())
    assert torch.allclose(g1, s1, rtol=1e-3, atol=tol), (g1 - s1).abs().max()
    assert torch.allclose(g2, s2, rtol=1e-3, atol=tol * 3), (g2 - s2).abs().max()


@pytest.mark.parametrize("N,h,w,Cc,pitch,coff", [(2, 8, 8, 64, 64, 0), (3, 6, 10, 128, 192, 64), (2, 3, 5, 512, 768, 256), (4, 24, 24, 32, 96, 32),
                                                    (2, 48, 48, 64, 192, 128),
                                                    # dense rows: the bulk-copy pipelined kernel (ring of whole-row stages)
                                                    (3, 48, 48, 64, 64, 0), (2, 24, 24, 128, 128, 0), (2, 6, 6, 512, 512, 0), (2, 17, 9, 32, 32, 0),
                                                    (1, 1, 1, 64, 64, 0), (2, 35, 3, 8, 8, 0), (1, 5, 124, 64, 64, 0), (1, 4, 200, 64, 64, 0)])
def test_up_transpose(N, h, w, Cc, pitch, coff):
    import ops
    gen = torch.Generator(device=DEV).manual_seed(17 + h)
    dx = torch.randn(N, 2 * h, 2 * w, pitch, generator=gen, device=DEV).to(torch.bfloat16)
    out = ops.up_transpose(dx, pitch, coff, N, h, w, Cc)
    a = torch.zeros(N, Cc, h, w, device=DEV, requires_grad=True)
    up = F.interpolate(a, scale_factor=2, mode="bilinear", align_corners=True)
    (up * dx[..., coff:coff + Cc].float().permute(0, 3, 1, 2)).sum().backward()
    torch.cuda.synchronize()
    err = (out.float().permute(0, 3, 1, 2) - a.grad).abs()
    assert bool((err <= 0.008 * a.grad.abs() + 1e-3 * a.grad.abs().max()).all()), err.max()


def test_bn_backward_reduce():
    import ops
    gen = torch.Generator(device=DEV).manual_seed(9)
    Cc, rows, M = 128, 53, 4096.0
    part = torch.randn(rows, 2, Cc, generator=gen, device=DEV)
    gamma = torch.randn(Cc, generator=gen, device=DEV)
    mean = torch.randn(Cc, generator=gen, device=DEV)
    invstd = torch.rand(Cc, generator=gen, device=DEV) + 0.5
    ws = ops.BnWorkspace(DEV)
    dg, db, k1, k2, k3 = (torch.empty(Cc, device=DEV) for _ in range(5))
    sums = torch.empty(2, Cc, dtype=torch.float64, device=DEV)
    ops.bn_backward_reduce(part, M, gamma, mean, invstd, ws, dg, db, k1, k2, k3, sums_out=sums)
    torch.cuda.synchronize()
    S1, S2 = part[:, 0].double().sum(0), part[:, 1].double().sum(0)
    g, r, mu = gamma.double(), invstd.double(), mean.double()
    assert torch.allclose(db.double(), S1, rtol=1e-5, atol=1e-5) and torch.allclose(dg.double(), S2, rtol=1e-5, atol=1e-5)
    assert torch.allclose(k1.double(), g * r, rtol=1e-5)
    assert torch.allclose(k2.double(), -g * r * r * S2 / M, rtol=1e-4, atol=1e-8)
    assert torch.allclose(k3.double(), -g * r * S1 / M + g * r * r * mu * S2 / M, rtol=1e-4, atol=1e-7)
    dg2, db2, k12, k22, k32 = (torch.empty(Cc, device=DEV) for _ in range(5))
    ops.bn_backward_from_sums(sums, M, gamma, mean, invstd, dg2, db2, k12, k22, k32)
    torch.cuda.synchronize()
    for a_, b_ in ((dg, dg2), (db, db2), (k1, k12), (k2, k22), (k3, k32)):
        assert torch.equal(a_, b_)


def _grad_operands(N, H, W, Cout, gen):
    gz = torch.randn(N, H, W, Cout, generator=gen, device=DEV).to(torch.bfloat16)
    y = torch.randn(N, H, W, Cout, generator=gen, device=DEV).to(torch.bfloat16)
    k1 = torch.randn(Cout, generator=gen, device=DEV)
    k2 = torch.randn(Cout, generator=gen, device=DEV) * 0.1
    k3 = torch.randn(Cout, generator=gen, device=DEV) * 0.1
    gy = (gz.float() * k1 + y.float() * k2 + k3).permute(0, 3, 1, 2)  # NCHW fp32
    return gz, y, k1, k2, k3, gy


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 16, 16, 32, 32), (2, 32, 32, 96, 32), (2, 16, 16, 192, 64), (2, 8, 8, 384, 128),
                                            (2, 6, 6, 768, 256), (2, 6, 6, 512, 512), (3, 12, 20, 160, 32), (16, 96, 96, 96, 32)])
def test_dgrad(N, H, W, Cin, Cout):
    import _lib as L
    import ops
    torch.backends.cudnn.allow_tf32 = False
    gen = torch.Generator(device=DEV).manual_seed(21 + Cin)
    gz, y, k1, k2, k3, gy = _grad_operands(N, H, W, Cout, gen)
    w = torch.randn(Cout, Cin, 3, 3, generator=gen, device=DEV) * (2.0 / (9 * Cout)) ** 0.5
    plan = ops.conv_plan(N, H, W, Cout, Cin)
    wp = ops.pack_weights(w, L.PACK_DGRAD, plan)
    src = ops.Source(L.SRC_GRAD, gz, Cout, k1, k3, data2=y, k2=k2)
    dx, _ = ops.conv3x3(N, H, W, [src], Cin, wp)
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(_bf(gy), _bf(w), padding=1)
    got = dx.float().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    assert bool((err <= 0.008 * ref.abs() + 2e-3 * ref.abs().max()).all()), err.max()
    # the engine's path: g_y materialised once (in place), dgrad reads it as a pre-activated source through cp.async
    gyb = ops.bn_bwd_apply(gz, y, k1, k2, k3, out=gz)
    dx2, _ = ops.conv3x3(N, H, W, [ops.Source(L.SRC_ACT, gyb, Cout)], Cin, wp)
    torch.cuda.synchronize()
    ref2 = F.conv_transpose2d(gyb.float().permute(0, 3, 1, 2), _bf(w), padding=1)
    err2 = (dx2.float().permute(0, 3, 1, 2) - ref2).abs()
    assert bool((err2 <= 0.008 * ref2.abs() + 2e-3 * ref2.abs().max()).all()), err2.max()


WG_CASES = [
    (2, 16, 16, [("same", 32)], 32),
    (2, 32, 32, [("same", 32), ("up", 64)], 32),
    (2, 16, 16, [("same", 64), ("same", 64), ("up", 128)], 64),
    (2, 8, 8, [("pool", 64)], 128),
    (2, 6, 6, [("same", 256), ("up", 512)], 256),
    (2, 6, 6, [("same", 512)], 512),
    (3, 12, 20, [("same", 32), ("same", 32), ("same", 32), ("same", 32), ("up", 64)], 32),
    (8, 96, 96, [("same", 32), ("up", 64)], 32),
    (2, 32, 32, [("same_act", 32), ("same_act", 32), ("up_act", 64)], 32),
    (2, 8, 8, [("pool_act", 64)], 128),
    (8, 96, 96, [("same_act", 32), ("up_act", 64)], 32),
    (2, 6, 6, [("same_act", 256), ("up_act", 512)], 256),
    # all sources pre-activated at the conv's resolution: the TMA-fed MN-major kernel (conv_wgrad_tma.cu)
    (2, 32, 32, [("same_act", 32), ("same_act", 32), ("same_act", 32)], 32),
    (8, 96, 96, [("same_act", 32), ("same_act", 64)], 32),
    (3, 48, 80, [("same_act", 32)] * 5, 32),
    (2, 16, 16, [("same_act", 64), ("same_act", 64), ("same_act", 128)], 64),
    (2, 24, 24, [("same_act", 128), ("same_act", 256)], 128),
    (3, 12, 20, [("same_act", 256), ("same_act", 512)], 256),
    (2, 6, 6, [("same_act", 512)], 512),
    (5, 6, 6, [("same_act", 256)], 512),
    # rows wider than a tensor-load box: column strips (X restricted to the strip row by row, gy with its real halo columns)
    (2, 5, 300, [("same_act", 64)], 32),
    (1, 4, 520, [("same_act", 32)], 64),              # Cin = 32: kernel rows stacked on the A side, 3 strips
    (2, 9, 301, [("same_act", 32), ("same_act", 128)], 64),
    (1, 3, 1000, [("same_act", 64)], 32),
]


@pytest.mark.parametrize("case", WG_CASES, ids=lambda c: "N%d_%dx%d_%s_%d" % (c[0], c[1], c[2], "+".join("%s%d" % s for s in c[3]), c[4]))
def test_wgrad(case):
    import ops
    from refops import source_tensor
    from test_conv_gpu import _mk_src
    torch.backends.cudnn.allow_tf32 = False
    N, H, W, srcs, Cout = case
    gen = torch.Generator(device=DEV).manual_seed(77 + Cout + H)
    sources, refs = [], []
    for kind, Cc in srcs:
        s, r = _mk_src(kind, N, H, W, Cc, gen, torch.device(DEV))
        sources.append(s)
        refs.append(r)
    cin = sum(c for _, c in srcs)
    gz, y, k1, k2, k3, gy = _grad_operands(N, H, W, Cout, gen)
    dw = torch.full((Cout, cin, 3, 3), float("nan"), device=DEV)
    gyb = ops.bn_bwd_apply(gz, y, k1, k2, k3)
    assert bool(((gyb.float().permute(0, 3, 1, 2) - gy).abs() <= 0.008 * gy.abs() + 1e-6).all())
    ops.conv3x3_wgrad(N, H, W, sources, Cout, gyb, dw)
    torch.cuda.synchronize()
    x = _bf(torch.cat([source_tensor(*r) for r in refs], 1))
    ref = torch.nn.grad.conv2d_weight(x, (Cout, cin, 3, 3), gyb.float().permute(0, 3, 1, 2), padding=1)
    err = (dw - ref).abs()
    assert bool(torch.isfinite(dw).all())
    assert bool((err <= 2e-3 * ref.abs() + 2e-3 * ref.abs().max()).all()), (err.max(), ref.abs().max())


def test_wgrad_stem():
    import _lib as L
    import ops
    N, H, W, Cin, Cout = 3, 32, 48, 3, 32
    gen = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn(N, Cin, H, W, generator=gen, device=DEV)
    gz, y, k1, k2, k3, gy = _grad_operands(N, H, W, Cout, gen)
    dw = torch.full((Cout, Cin, 3, 3), float("nan"), device=DEV)
    gyb = ops.bn_bwd_apply(gz, y, k1, k2, k3)
    ops.conv3x3_wgrad(N, H, W, [ops.Source(L.SRC_INPUT, x, Cin)], Cout, gyb, dw)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(_bf(x), (Cout, Cin, 3, 3), gyb.float().permute(0, 3, 1, 2), padding=1)
    err = (dw - ref).abs()
    assert bool((err <= 2e-3 * ref.abs() + 2e-3 * ref.abs().max()).all()), (err.max(), ref.abs().max())
