"""Parity of the fused implicit-GEMM conv (forward + BN statistics) against a torch fp32 reference on the device."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk_src(kind, N, H, W, C, gen, dev):
    """Random raw producer output at the resolution the kind implies + random BN coefficients (both signs)."""
    import _lib as L
    import ops
    base = kind.replace("_act", "")
    hs, ws = {"same": (H, W), "pool": (2 * H, 2 * W), "up": (H // 2, W // 2)}[base]
    y = (torch.randn(N, hs, ws, C, generator=gen, device=dev) * 1.5).to(torch.bfloat16)
    scale = torch.randn(C, generator=gen, device=dev)
    shift = torch.randn(C, generator=gen, device=dev) * 0.5
    if kind.endswith("_act"):
        # pre-activated source: the loader reads a = relu(scale*y+shift) materialised by nunet_bn_apply
        a = ops.bn_apply(y, scale, shift)
        ref_a = torch.relu(y.float() * scale + shift)
        assert bool(((a.float() - ref_a).abs() <= 0.008 * ref_a.abs() + 1e-6).all())   # fma vs mul+add: at most one bf16 ulp
        k = {"same": L.SRC_ACT, "pool": L.SRC_POOL_ACT, "up": L.SRC_UP_ACT}[base]
        one, zero = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        return ops.Source(k, a, C), (base, a, one, zero)
    k = {"same": L.SRC_SAME, "pool": L.SRC_POOL, "up": L.SRC_UP}[kind]
    return ops.Source(k, y, C, scale, shift), (kind, y, scale, shift)


CASES = [
    # (N, H, W, [(kind, C), ...], Cout)
    (2, 32, 32, [("same", 32)], 32),
    (3, 16, 16, [("same", 64)], 64),
    (4, 8, 8, [("same", 128)], 128),
    (5, 6, 6, [("same", 256)], 256),
    (3, 6, 6, [("same", 512)], 512),
    (2, 16, 16, [("pool", 32)], 64),
    (2, 6, 6, [("pool", 256)], 512),
    (2, 32, 32, [("same", 32), ("up", 64)], 32),
    (2, 32, 32, [("same", 32), ("same", 32), ("same", 32), ("same", 32), ("up", 64)], 32),
    (2, 12, 12, [("same", 256), ("up", 512)], 256),
    (2, 16, 16, [("same", 64), ("same", 64), ("same", 64), ("up", 128)], 64),
    (3, 48, 80, [("same", 32), ("up", 64)], 32),
    (16, 96, 96, [("same", 32)], 32),
    (16, 48, 48, [("same", 64), ("up", 128)], 64),
    (2, 32, 32, [("same_act", 32)], 32),
    (2, 16, 16, [("pool_act", 32)], 64),
    (2, 32, 32, [("same_act", 32), ("same_act", 32), ("same_act", 32), ("same_act", 32), ("up_act", 64)], 32),
    (3, 12, 20, [("same_act", 256), ("up_act", 512)], 256),
    (16, 96, 96, [("same_act", 32), ("up_act", 64)], 32),
    (2, 6, 6, [("up_act", 512)], 256),
    # rows wider than a tensor-load box: column strips (2, 3 and 4 strips; widths that do not divide evenly)
    (2, 5, 301, [("same_act", 32)], 32),
    (1, 4, 520, [("same_act", 64), ("same_act", 32)], 64),
    (2, 3, 256, [("same_act", 32)], 96),
    (1, 2, 1000, [("same_act", 32)], 32),
    (2, 7, 21, [("same_act", 32)], 32),     # odd width: stage rows are only 64-byte aligned -> cp.async copy mode
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "N%d_%dx%d_%s_%d" % (c[0], c[1], c[2], "+".join("%s%d" % s for s in c[3]), c[4]))
def test_conv_fwd_sources(case):
    import _lib as L
    import ops
    from refops import conv_ref, source_tensor
    N, H, W, srcs, Cout = case
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(1234 + N + H + Cout)
    sources, refs = [], []
    for kind, Cc in srcs:
        s, r = _mk_src(kind, N, H, W, Cc, gen, dev)
        sources.append(s)
        refs.append(r)
    cin = sum(c for _, c in srcs)
    w = torch.randn(Cout, cin, 3, 3, generator=gen, device=dev) * (2.0 / (9 * cin)) ** 0.5
    plan = ops.conv_plan(N, H, W, cin, Cout)
    wp = ops.pack_weights(w, L.PACK_FWD, plan)
    out, part = ops.conv3x3(N, H, W, sources, Cout, wp, want_stats=True)
    torch.cuda.synchronize()
    x = torch.cat([source_tensor(*r) for r in refs], 1)
    ref = conv_ref(x, w)
    got = out.float().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    tol = 0.008 * ref.abs() + 2e-3 * ref.abs().max()
    assert bool((err <= tol).all()), "max err %g (ref max %g), T=%d NT=%d CK=%d S=%d" % (err.max(), ref.abs().max(), plan.T, plan.NT, plan.CK, plan.S)
    # the statistics are those of the STORED bf16 values (what the BatchNorm that follows normalises): tight against the output itself,
    # and within bf16 rounding of the fp32 reference's sums
    s = part[:, 0, :].double().sum(0)
    ss = part[:, 1, :].double().sum(0)
    os_, oss = got.double().sum((0, 2, 3)), (got.double() ** 2).sum((0, 2, 3))
    assert torch.allclose(s, os_, rtol=1e-4, atol=1e-4 * float(oss.max().sqrt())), (s - os_).abs().max()
    assert torch.allclose(ss, oss, rtol=1e-4), ((ss - oss) / oss).abs().max()
    rs = ref.double().sum((0, 2, 3))
    rss = (ref.double() ** 2).sum((0, 2, 3))
    assert torch.allclose(s, rs, rtol=4e-3, atol=4e-3 * float(rss.max().sqrt())), (s - rs).abs().max()
    assert torch.allclose(ss, rss, rtol=4e-3), ((ss - rss) / rss).abs().max()


@pytest.mark.parametrize("cin_real", [3, 1, 4])
def test_conv_fwd_stem_input(cin_real):
    import _lib as L
    import ops
    from refops import conv_ref
    N, H, W, Cout = 3, 32, 48, 32
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(7)
    x = torch.randn(N, cin_real, H, W, generator=gen, device=dev)
    w = torch.randn(Cout, cin_real, 3, 3, generator=gen, device=dev) * 0.2
    plan = ops.conv_plan(N, H, W, 16, Cout)
    wp = ops.pack_weights(w, L.PACK_FWD, plan)
    out, part = ops.conv3x3(N, H, W, [ops.Source(L.SRC_INPUT, x, cin_real)], Cout, wp, want_stats=True)
    torch.cuda.synchronize()
    ref = conv_ref(x, w)
    got = out.float().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    assert bool((err <= 0.008 * ref.abs() + 2e-3 * ref.abs().max()).all()), err.max()


def test_bn_finalize_matches_batchnorm():
    import ops
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(3)
    Cc, rows = 64, 77
    part = torch.rand(rows, 2, Cc, generator=gen, device=dev)
    part[:, 1, :] += 3.0
    count = 1000.0
    gamma = torch.randn(Cc, generator=gen, device=dev)
    beta = torch.randn(Cc, generator=gen, device=dev)
    bias = torch.randn(Cc, generator=gen, device=dev)
    rm = torch.randn(Cc, generator=gen, device=dev)
    rv = torch.rand(Cc, generator=gen, device=dev) + 0.5
    rm0, rv0 = rm.clone(), rv.clone()
    ws = ops.BnWorkspace(dev)
    scale, shift, mean, invstd = (torch.empty(Cc, device=dev) for _ in range(4))
    sums = torch.empty(2, Cc, dtype=torch.float64, device=dev)
    for _ in range(2):  # twice: the counters must self-reset
        rm.copy_(rm0)
        rv.copy_(rv0)
        ops.bn_finalize_train(part, count, gamma, beta, bias, rm, rv, 0.1, 1e-5, ws, scale, shift, mean, invstd, sums_out=sums)
    torch.cuda.synchronize()
    S = part[:, 0].double().sum(0)
    SS = part[:, 1].double().sum(0)
    m = S / count
    v = SS / count - m * m
    assert torch.allclose(sums[0], S, rtol=1e-6) and torch.allclose(sums[1], SS, rtol=1e-6)
    assert torch.allclose(mean.double(), m, rtol=1e-5, atol=1e-6)
    inv = 1.0 / torch.sqrt(v + 1e-5)
    assert torch.allclose(invstd.double(), inv, rtol=1e-5)
    assert torch.allclose(scale.double(), gamma.double() * inv, rtol=1e-5, atol=1e-6)
    assert torch.allclose(shift.double(), beta.double() - m * gamma.double() * inv, rtol=1e-4, atol=1e-5)
    assert torch.allclose(rm.double(), 0.9 * rm0.double() + 0.1 * (m + bias.double()), rtol=1e-5, atol=1e-6)
    assert torch.allclose(rv.double(), 0.9 * rv0.double() + 0.1 * v * count / (count - 1), rtol=1e-5, atol=1e-6)
    # SyncBN path: sums -> finalize
    scale2, shift2 = torch.empty_like(scale), torch.empty_like(shift)
    ops.bn_finalize_from_sums(sums, count, gamma, beta, None, None, None, 0.1, 1e-5, scale2, shift2, None, None)
    torch.cuda.synchronize()
    assert torch.allclose(scale2, scale) and torch.allclose(shift2, shift)


def test_upsample2x_matches_torch():
    import ops
    import torch.nn.functional as F
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(5)
    for (N, h, w, Cc) in [(2, 6, 6, 512), (3, 12, 20, 64), (2, 48, 48, 64), (1, 3, 5, 128)]:
        a = torch.rand(N, h, w, Cc, generator=gen, device=dev).to(torch.bfloat16)
        up = ops.upsample2x(a)
        ref = F.interpolate(a.float().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
        torch.cuda.synchronize()
        assert bool(((up.float() - ref).abs() <= 0.008 * ref.abs() + 1e-6).all())


C32_CASES = [(2, 16, 16, [64]), (2, 32, 32, [32, 64]), (3, 48, 80, [32, 32, 64]), (4, 96, 96, [32] * 4 + [64]), (2, 6, 6, [64]), (5, 96, 96, [32, 32]),
             (1, 16, 240, [32, 32, 32]),
             (2, 5, 300, [64]), (1, 4, 521, [32, 64])]      # column strips (2 and 3; last strip narrower)


@pytest.mark.parametrize("case", C32_CASES, ids=lambda c: "N%d_%dx%d_cin%d" % (c[0], c[1], c[2], sum(c[3])))
@pytest.mark.parametrize("mode", ["fwd", "dgrad"])
def test_stacked_cout32_kernel_matches_torch_and_the_generic_kernel(case, mode):
    """conv_c32.cu (kernel columns stacked into N = 96, column shift applied to the accumulators) against torch on identical
    bf16 operands and against the generic implicit-GEMM kernel: outputs equal up to the bf16 rounding of slightly different fp32
    accumulation orders, BatchNorm partial sums to 2e-5 of the sum of magnitudes."""
    import torch.nn.functional as F
    import _lib as L
    import ops
    dev = "cuda:0"
    N, H, W, Cs = case
    gen = torch.Generator(device=dev).manual_seed(7)
    cin = sum(Cs)
    srcs, xs = [], []
    for c in Cs:
        a = (torch.rand(N, H, W, c, generator=gen, device=dev) - 0.3).to(torch.bfloat16)
        srcs.append(ops.Source(L.SRC_ACT, a, c))
        xs.append(a.float().permute(0, 3, 1, 2))
    assert ops.c32_eligible(N, H, W, cin, 32)
    plan = ops.conv_plan(N, H, W, cin, 32)
    if mode == "fwd":
        w = torch.randn(32, cin, 3, 3, generator=gen, device=dev) * 0.1
        wp, wp0 = ops.pack_weights(w, L.PACK_FWD_C32, plan), ops.pack_weights(w, L.PACK_FWD, plan)
        ref = F.conv2d(torch.cat(xs, 1), w.to(torch.bfloat16).float(), padding=1)
    else:
        w = torch.randn(cin, 32, 3, 3, generator=gen, device=dev) * 0.1
        wp, wp0 = ops.pack_weights(w, L.PACK_DGRAD_C32, plan), ops.pack_weights(w, L.PACK_DGRAD, plan)
        ref = F.conv_transpose2d(torch.cat(xs, 1), w.to(torch.bfloat16).float(), padding=1)
    y, part = ops.conv3x3(N, H, W, srcs, 32, wp, want_stats=True, stacked=True)
    y0, _ = ops.conv3x3(N, H, W, srcs, 32, wp0, want_stats=True)
    y_plain, none = ops.conv3x3(N, H, W, srcs, 32, wp, want_stats=False, stacked=True)
    torch.cuda.synchronize()
    refn = ref.permute(0, 2, 3, 1)
    tol = 0.006 * float(refn.abs().max()) + 1e-3          # half a bf16 ulp of the largest output + accumulation-order noise
    assert float((y.float() - refn).abs().max()) <= tol
    assert float((y.float() - y0.float()).abs().max()) <= tol
    assert torch.equal(y, y_plain) and none is None
    flat = refn.reshape(-1, 32).double()
    s_ref, q_ref = flat.sum(0), (flat ** 2).sum(0)
    s, q = part[:, 0].double().sum(0), part[:, 1].double().sum(0)
    # fp32 partial sums: the error scales with the sum of magnitudes (the per-channel sums themselves may cancel to ~0);
    # the bf16-rounded outputs are what the reference sums, the kernel sums its fp32 accumulators (BatchNorm statistics)
    l1 = flat.abs().sum(0)
    assert float(((s - s_ref).abs() / l1).max()) <= 2e-5
    assert float(((q - q_ref).abs() / (q_ref.abs() + 1.0)).max()) <= 1e-4
