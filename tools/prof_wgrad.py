"""Per-role cycle breakdown of the wgrad kernel (build with NUNET_NVCC_EXTRA=-DNUNET_PROFILE_WAITS)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import ops, _lib as L
lib = L.lib()
rd = lib.nunet_debug_read_wgrad_prof
rd.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_ulonglong * 16)()
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
CASES = [(256, 96, 96, [("same", 32)], 32), (256, 96, 96, [("act", 32)] * 4 + [("act", 64)], 32),
         (256, 48, 48, [("same", 64)], 64), (256, 48, 48, [("act", 64)] * 3 + [("act", 128)], 64), (256, 24, 24, [("same", 128)], 128),
         (256, 24, 24, [("act", 128)] * 2 + [("act", 256)], 128),
         (256, 12, 12, [("act", 256), ("act", 512)], 256), (256, 6, 6, [("same", 512)], 512)]
for N, H, W, srcs, Cout in CASES:
    sources = []
    for k, Cc in srcs:
        hs, ws = (H // 2, W // 2) if k.startswith("up") else (H, W)
        y = torch.randn(N, hs, ws, Cc, generator=g, device=dev).to(torch.bfloat16)
        sc = torch.randn(Cc, generator=g, device=dev); sh = torch.randn(Cc, generator=g, device=dev)
        kind = {"act": L.SRC_ACT, "same": L.SRC_SAME, "up_act": L.SRC_UP_ACT}[k]
        sources.append(ops.Source(kind, y, Cc, sc, sh))
    cin = sum(c for _, c in srcs)
    plan = ops.wgrad_plan(N, H, W, cin, Cout)
    gy = torch.randn(N, H, W, Cout, generator=g, device=dev).to(torch.bfloat16)
    dw = torch.empty(Cout, cin, 3, 3, device=dev)
    for it in range(2):
        torch.cuda.synchronize(); rd(buf, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.conv3x3_wgrad(N, H, W, sources, Cout, gy, dw); e1.record()
        torch.cuda.synchronize(); rd(buf, 0)
    nct = plan.splits * plan.m_tiles * plan.n_tiles
    v = [x / nct / 1e3 for x in buf]
    ms = e0.elapsed_time(e1); fl = 2.0 * N * H * W * cin * Cout * 9
    print("wgrad %dx%d cin %d cout %d: %.3f ms %.0f TF/s | KP=%d S=%d cpc=%d ctas=%d (%.1f/SM) m_tiles=%d n_tiles=%d" % (H, W, cin, Cout, ms, fl / ms / 1e9, plan.KP, plan.S, plan.cpc, nct, nct / 148.0, plan.m_tiles, plan.n_tiles))
    print("   kcycles/CTA producer: wait_empty %.0f table %.0f gy %.0f x %.0f wait_copy %.0f | mma: wait_full %.0f issue %.0f" % (v[0], v[1], v[2], v[3], v[4], v[8], v[9]))
