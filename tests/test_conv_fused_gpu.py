"""Round-2 fusions of the conv kernel, each against the unfused launches it replaces (bitwise or to fp32 round-off):
  * BatchNorm finalisation in the conv's tail (last CTA) == nunet_bn_finalize_train on the same partial sums
  * dgrad with one dense output tensor per concatenated source == channel slices of the single wide output
  * dgrad of a block's second conv with the block-internal gradient router fused into its epilogue == dgrad -> nunet_grad_router
    -> nunet_bn_backward_reduce, including the BatchNorm-backward coefficients from the kernel's tail
  * the gradient router's own tail == nunet_bn_backward_reduce
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bf(gen, *shape, scale=1.0):
    return (torch.randn(*shape, generator=gen, device=DEV) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("N,H,W,cin,cout,stacked", [(2, 32, 32, 32, 32, False), (16, 96, 96, 32, 32, False), (3, 16, 16, 64, 64, False),
                                                     (2, 12, 12, 256, 256, False), (3, 6, 6, 512, 512, False), (4, 48, 80, 96, 32, True),
                                                     (2, 32, 272, 32, 32, False)])
def test_bn_tail_equals_separate_finalize(N, H, W, cin, cout, stacked):
    import _lib as L
    import ops
    gen = torch.Generator(device=DEV).manual_seed(7 + cin + cout)
    a = _bf(gen, N, H, W, cin).abs()
    w = torch.randn(cout, cin, 3, 3, generator=gen, device=DEV) * 0.05
    gamma, beta, bias = (torch.randn(cout, generator=gen, device=DEV) for _ in range(3))
    plan = ops.conv_plan(N, H, W, cin, cout)
    stacked = stacked and ops.c32_eligible(N, H, W, cin, cout)
    wp = ops.pack_weights(w, L.PACK_FWD_C32 if stacked else L.PACK_FWD, plan)
    src = [ops.Source(L.SRC_ACT, a, cin)]
    f32 = lambda: torch.empty(cout, device=DEV)
    # reference: statistics epilogue + separate finalize launch
    y0, part = ops.conv3x3(N, H, W, src, cout, wp, want_stats=True, stacked=stacked)
    rm0, rv0 = torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV)
    sc0, sh0, mu0, is0 = f32(), f32(), f32(), f32()
    ops.bn_finalize_train(part, float(N * H * W), gamma, beta, bias, rm0, rv0, 0.1, 1e-5, ops.BnWorkspace(DEV), sc0, sh0, mu0, is0)
    # fused: the conv's last CTA finalises
    rm1, rv1 = torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV)
    sc1, sh1, mu1, is1 = f32(), f32(), f32(), f32()
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    for rep in range(3):      # the counter resets itself: repeated launches keep working
        rm1.zero_(); rv1.fill_(1.0)
        tail = ops.bn_tail(counter, float(N * H * W), gamma, beta, bias, rm1, rv1, 0.1, 1e-5, sc1, sh1, mu1, is1)
        y1, _ = ops.conv3x3(N, H, W, src, cout, wp, want_stats=True, stacked=stacked, tail=tail)
        torch.cuda.synchronize()
        assert int(counter) == 0
        assert torch.equal(y0, y1)
        # same partial rows, both reduced in fp64 (different association): agreement to fp32 round-off
        for a_, b_ in ((sc0, sc1), (sh0, sh1), (mu0, mu1), (is0, is1), (rm0, rm1), (rv0, rv1)):
            assert torch.allclose(a_, b_, rtol=2e-6, atol=1e-6), float((a_ - b_).abs().max())


@pytest.mark.parametrize("N,H,W,cout,segs", [(2, 32, 32, 32, [32, 64]), (3, 48, 80, 32, [32, 32, 32, 32, 64]), (2, 16, 16, 64, [64, 64, 128]),
                                             (2, 12, 12, 256, [256, 512]), (2, 8, 8, 128, [128, 128, 256]), (2, 32, 272, 32, [32, 64])])
def test_dgrad_per_source_outputs_equal_slices_of_the_wide_output(N, H, W, cout, segs):
    import _lib as L
    import ops
    gen = torch.Generator(device=DEV).manual_seed(11 + sum(segs))
    cin = sum(segs)
    gy = _bf(gen, N, H, W, cout)
    w = torch.randn(cout, cin, 3, 3, generator=gen, device=DEV) * 0.05
    plan = ops.conv_plan(N, H, W, cout, cin)
    wpd = ops.pack_weights(w, L.PACK_DGRAD, plan)
    src = [ops.Source(L.SRC_ACT, gy, cout)]
    wide, _ = ops.conv3x3(N, H, W, src, cin, wpd, cls="conv_dgrad")
    parts, _ = ops.conv3x3(N, H, W, src, cin, wpd, cls="conv_dgrad", out_segs=segs)
    torch.cuda.synchronize()
    off = 0
    for t, c in zip(parts, segs):
        assert t.shape == (N, H, W, c)
        assert torch.equal(t, wide[..., off:off + c])
        off += c


@pytest.mark.parametrize("N,H,W,cmid,cout,stacked", [(2, 32, 32, 32, 32, False), (16, 96, 96, 32, 32, False), (3, 16, 24, 64, 64, False), (2, 12, 12, 256, 256, False),
                                                     (3, 6, 6, 512, 512, False), (2, 24, 24, 128, 128, False), (2, 32, 272, 32, 32, False),
                                                     (2, 32, 32, 32, 64, True), (16, 96, 96, 32, 96, True), (3, 48, 80, 32, 64, True)])
def test_fused_block_router_equals_dgrad_then_router(N, H, W, cmid, cout, stacked):
    """conv2's dgrad with mask = (y1, scale1, shift1, mean1, invstd1) must produce g_z of conv1 and its BatchNorm-backward
    coefficients exactly as the separate launches do (the fused path masks the fp32 accumulator, the separate one the bf16 dX:
    g_z agrees to the bf16 rounding of dX, i.e. bitwise after the same rounding; sums to fp32 round-off)."""
    import _lib as L
    import ops
    gen = torch.Generator(device=DEV).manual_seed(5 + cmid + H)
    gy2 = _bf(gen, N, H, W, cout)
    y1 = _bf(gen, N, H, W, cmid, scale=1.5)
    w2 = torch.randn(cout, cmid, 3, 3, generator=gen, device=DEV) * 0.05
    sc, sh = torch.randn(cmid, generator=gen, device=DEV), torch.randn(cmid, generator=gen, device=DEV) * 0.5
    mu, rs = torch.randn(cmid, generator=gen, device=DEV) * 0.3, torch.rand(cmid, generator=gen, device=DEV) + 0.5
    gamma = torch.randn(cmid, generator=gen, device=DEV)
    plan = ops.conv_plan(N, H, W, cout, cmid)
    wpd0 = ops.pack_weights(w2, L.PACK_DGRAD, plan)
    stacked = stacked and ops.c32_eligible(N, H, W, cout, cmid)        # the stacked Cout = 32 kernel's epilogue has the same fusion
    wpd = ops.pack_weights(w2, L.PACK_DGRAD_C32, plan) if stacked else wpd0
    src = [ops.Source(L.SRC_ACT, gy2, cout)]
    count = float(N * H * W)
    f32 = lambda: torch.empty(cmid, device=DEV)
    # separate launches
    dx, _ = ops.conv3x3(N, H, W, src, cmid, wpd0, cls="conv_dgrad")
    gz0, part0 = ops.grad_router(N, H, W, cmid, y1, sc, sh, mu, rs, [ops.Consumer(dx, cmid, 0, L.SRC_SAME)])
    dg0, db0, k10, k20, k30 = f32(), f32(), f32(), f32(), f32()
    ops.bn_backward_reduce(part0, count, gamma, mu, rs, ops.BnWorkspace(DEV), dg0, db0, k10, k20, k30)
    # fused
    dg1, db1, k11, k21, k31 = f32(), f32(), f32(), f32(), f32()
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    bt = ops.bn_bwd_tail(counter, count, gamma, mu, rs, dg1, db1, k11, k21, k31)
    gz1, _ = ops.conv3x3(N, H, W, src, cmid, wpd, cls="conv_dgrad", mask=ops.Mask(y1, sc, sh, mu, rs), bwd_tail=bt, stacked=stacked)
    torch.cuda.synchronize()
    assert int(counter) == 0
    if stacked:   # the stacked kernel adds the three kernel-column partial sums in another order: fp32 round-off before the bf16 rounding
        assert bool(((gz0.float() - gz1.float()).abs() <= 0.008 * gz0.float().abs() + 1e-5).all())
    else:
        assert torch.equal(gz0, gz1)
    # the separate path sums bf16-rounded g_z, the fused one the fp32 accumulators: relative to the per-channel scale of the sums
    for a_, b_, name in ((dg0, dg1, "dgamma"), (db0, db1, "dbeta"), (k20, k21, "k2"), (k30, k31, "k3")):
        tol = 4e-3 * float(a_.abs().max()) + 1e-6
        assert float((a_ - b_).abs().max()) <= tol, (name, float((a_ - b_).abs().max()), tol)
    assert torch.equal(k10, k11)


@pytest.mark.parametrize("N,H,W,C,ncons", [(2, 32, 32, 32, 3), (3, 16, 16, 64, 1), (2, 12, 12, 256, 2), (3, 6, 6, 512, 1)])
def test_router_tail_equals_separate_reduce(N, H, W, C, ncons):
    import _lib as L
    import ops
    gen = torch.Generator(device=DEV).manual_seed(3 + C)
    y = _bf(gen, N, H, W, C, scale=1.5)
    sc, sh = torch.randn(C, generator=gen, device=DEV), torch.randn(C, generator=gen, device=DEV) * 0.5
    mu, rs = torch.randn(C, generator=gen, device=DEV) * 0.3, torch.rand(C, generator=gen, device=DEV) + 0.5
    gamma = torch.randn(C, generator=gen, device=DEV)
    cons = [ops.Consumer(_bf(gen, N, H, W, C), C, 0, L.SRC_SAME) for _ in range(ncons)]
    count = float(N * H * W)
    f32 = lambda: torch.empty(C, device=DEV)
    gz0, part0 = ops.grad_router(N, H, W, C, y, sc, sh, mu, rs, cons)
    ref = [f32() for _ in range(5)]
    ops.bn_backward_reduce(part0, count, gamma, mu, rs, ops.BnWorkspace(DEV), *ref)
    got = [f32() for _ in range(5)]
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    gz1, _ = ops.grad_router(N, H, W, C, y, sc, sh, mu, rs, cons, tail=ops.bn_bwd_tail(counter, count, gamma, mu, rs, *got))
    torch.cuda.synchronize()
    assert int(counter) == 0 and torch.equal(gz0, gz1)
    for a_, b_ in zip(ref, got):
        assert torch.allclose(a_, b_, rtol=2e-6, atol=1e-7 * max(1.0, float(a_.abs().max()))), float((a_ - b_).abs().max())


@pytest.mark.parametrize("N,H,W,C,a,pool,up,K", [(3, 32, 32, 32, True, False, False, 0), (3, 32, 32, 32, True, False, False, 1), (2, 48, 80, 32, False, False, False, 4),
                                                (2, 32, 48, 32, True, True, False, 0), (2, 16, 24, 64, True, True, True, 0), (3, 8, 8, 128, True, True, True, 0),
                                                (2, 6, 10, 256, True, True, True, 0), (5, 2, 2, 512, False, False, True, 0), (2, 6, 6, 512, False, False, True, 0),
                                                (2, 48, 48, 64, False, False, True, 0), (2, 20, 12, 64, True, False, True, 0), (2, 16, 136, 64, True, True, True, 0)])
def test_activate_equals_separate_kernels_and_torch(N, H, W, C, a, pool, up, K):
    """nunet_bn_activate against the four launches it replaces (bn_apply, maxpool2x2, upsample2x, head_fwd) and torch fp32."""
    import ops
    import torch.nn.functional as F
    gen = torch.Generator(device=DEV).manual_seed(17 + H + C + K)
    y = _bf(gen, N, H, W, C, scale=1.5)
    sc, sh = torch.randn(C, generator=gen, device=DEV), torch.randn(C, generator=gen, device=DEV) * 0.5
    head = None
    if K:
        wh, bh = torch.randn(K, 32, generator=gen, device=DEV) * 0.3, torch.randn(K, generator=gen, device=DEV)
        head = (wh, bh)
    ga, gp, gu, gl = ops.bn_activate(y, sc, sh, a, pool, up, head)
    ref_a = ops.bn_apply(y, sc, sh)
    torch.cuda.synchronize()
    assert (ga is not None) == a and (gp is not None) == pool and (gu is not None) == up and (gl is not None) == bool(K)
    if a:
        assert torch.equal(ga, ref_a)
    if pool:
        assert torch.equal(gp, ops.maxpool2x2(ref_a))
    if up:
        # same arithmetic as nunet_upsample2x (blend of the bf16 activation in fp32, one rounding); fma contraction may differ by an ulp
        ref_u = ops.upsample2x(ref_a).float()
        assert float((gu.float() - ref_u).abs().max()) <= 0.008 * float(ref_u.abs().max()) + 1e-6
        t_u = F.interpolate(ref_a.float().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
        assert bool(((gu.float() - t_u).abs() <= 0.008 * t_u.abs() + 1e-5).all())
    if K:
        ref_l = ops.head_fwd(y, sc, sh, wh, bh, N, H, W)
        a32 = torch.relu(y.float() * sc + sh)
        t_l = (a32.reshape(-1, 32) @ wh.t() + bh).reshape(N, H, W, K).permute(0, 3, 1, 2)
        assert torch.allclose(gl, ref_l, rtol=1e-5, atol=1e-5) and torch.allclose(gl, t_l, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("N,H,W,C,ns,pool,head", [(2, 32, 32, 32, 4, True, False), (3, 16, 24, 64, 4, True, False), (2, 8, 8, 128, 3, True, False),
                                                  (2, 12, 12, 256, 2, True, False), (2, 32, 32, 32, 0, False, True), (3, 48, 80, 32, 3, False, True),
                                                  (2, 32, 32, 32, 1, False, True), (2, 16, 16, 64, 2, False, False), (2, 32, 272, 32, 2, False, True)])
def test_router_v2_equals_v1_and_separate_head_gradients(N, H, W, C, ns, pool, head):
    """The window-per-thread / fused-head router (grad_router_v2_kernel) against round 1's kernels (NUNET_ROUTER_V1 is read once per
    process, so the v1 result comes from the generic path: a POOL consumer together with a UP-free list is served by both) and, for the
    head, against nunet_head_bwd_params."""
    import _lib as L
    import ops
    gen = torch.Generator(device=DEV).manual_seed(23 + C + ns)
    y = _bf(gen, N, H, W, C, scale=1.5)
    sc, sh = torch.randn(C, generator=gen, device=DEV), torch.randn(C, generator=gen, device=DEV) * 0.5
    mu, rs = torch.randn(C, generator=gen, device=DEV) * 0.3, torch.rand(C, generator=gen, device=DEV) + 0.5
    gamma = torch.randn(C, generator=gen, device=DEV)
    cons = [ops.Consumer(_bf(gen, N, H, W, C), C, 0, L.SRC_SAME) for _ in range(ns)]
    if pool:
        cons.append(ops.Consumer(_bf(gen, N, H // 2, W // 2, C), C, 0, L.SRC_POOL))
    dl = wh = None
    if head:
        dl = torch.randn(N, 1, H, W, generator=gen, device=DEV) * 0.1
        wh = torch.randn(1, C, generator=gen, device=DEV) * 0.3
    # torch reference of the routing: g_a = sum same + max-pool backward (first maximum) + dl*wh ; g_z = g_a*[a>0]
    a = torch.relu(y.float() * sc + sh)
    ga = torch.zeros_like(a)
    for c in cons[:ns]:
        ga += c.dx.float()
    if pool:
        an = a.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        torch.nn.functional.max_pool2d(an, 2, 2).backward(cons[-1].dx.float().permute(0, 3, 1, 2).contiguous())
        ga += an.grad.permute(0, 2, 3, 1)
    if head:
        ga += dl[:, 0, :, :, None] * wh[0]
    z = y.float() * sc + sh
    gz_ref = torch.where(z > 0, ga, torch.zeros_like(ga))
    count = float(N * H * W)
    got = [torch.empty(C, device=DEV) for _ in range(5)]
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    hg = (torch.empty(1, C, device=DEV), torch.empty(1, device=DEV)) if head else None
    gz, _ = ops.grad_router(N, H, W, C, y, sc, sh, mu, rs, cons, dlogits=dl, wh=wh, tail=ops.bn_bwd_tail(counter, count, gamma, mu, rs, *got), head_grads=hg)
    torch.cuda.synchronize()
    assert int(counter) == 0
    assert bool(((gz.float() - gz_ref).abs() <= 0.008 * gz_ref.abs() + 1e-6).all())
    S1 = gz_ref.double().sum((0, 1, 2))
    S2 = (gz_ref.double() * ((y.double() - mu.double()) * rs.double())).sum((0, 1, 2))
    scale1 = float(gz_ref.abs().double().sum((0, 1, 2)).max())
    assert float((got[1].double() - S1).abs().max()) <= 2e-3 * scale1 + 1e-6      # dbeta  (g_z is bf16-rounded on the way)
    assert float((got[0].double() - S2).abs().max()) <= 4e-3 * scale1 * 3 + 1e-6  # dgamma
    if head:
        dwh_ref, dbh_ref = ops.head_bwd_params(y, sc, sh, dl, N, H, W, 1)
        torch.cuda.synchronize()
        assert torch.allclose(hg[0], dwh_ref, rtol=1e-4, atol=1e-5 * float(dwh_ref.abs().max()) + 1e-7)
        assert torch.allclose(hg[1], dbh_ref, rtol=1e-4, atol=1e-6)
