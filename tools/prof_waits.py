"""Per-role wait/work cycle breakdown of the conv kernel (library must be built with NUNET_NVCC_EXTRA=-DNUNET_PROFILE_WAITS)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import ops, _lib as L
lib = L.lib()
rd = lib.nunet_debug_read_wait_prof
rd.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_ulonglong * 16)()
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
CASES = [("fwd", 256, 96, 96, [("act", 32)], 32), ("dgrad", 256, 96, 96, [("act", 32)], 32), ("mask", 256, 96, 96, [("act", 32)], 32), ("mask", 256, 48, 48, [("act", 64)], 64), ("dgrad", 256, 96, 96, [("act", 32)], 192),
         ("dgrad", 256, 96, 96, [("act", 32)], 96), ("fwd", 256, 48, 48, [("act", 64)], 64), ("fwd", 256, 24, 24, [("act", 128)], 128), ("fwd", 256, 12, 12, [("act", 256)], 256), ("dgrad", 256, 48, 48, [("act", 64)], 320)]
for kind, N, H, W, srcs, Cout in CASES:
    sources = []
    for k, Cc in srcs:
        hs, ws = {"same": (H, W), "grad": (H, W), "act": (H, W), "up": (H // 2, W // 2)}[k]
        y = torch.randn(N, hs, ws, Cc, generator=g, device=dev).to(torch.bfloat16)
        sc = torch.randn(Cc, generator=g, device=dev); sh = torch.randn(Cc, generator=g, device=dev)
        if k == "grad":
            sources.append(ops.Source(L.SRC_GRAD, y, Cc, sc, sh, data2=y, k2=sc))
        else:
            sources.append(ops.Source({"same": L.SRC_SAME, "up": L.SRC_UP, "act": L.SRC_ACT}[k], y, Cc, sc, sh))
    cin = sum(c for _, c in srcs)
    plan = ops.conv_plan(N, H, W, cin, Cout)
    w = torch.randn(Cout, cin, 3, 3, generator=g, device=dev) * 0.05 if kind == "fwd" else torch.randn(cin, Cout, 3, 3, generator=g, device=dev) * 0.05
    wp = ops.pack_weights(w, L.PACK_FWD if kind == "fwd" else L.PACK_DGRAD, plan)
    for it in range(2):
        torch.cuda.synchronize(); rd(buf, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        mk = None
        if kind == "mask":
            y1 = torch.randn(N, H, W, Cout, generator=g, device=dev).to(torch.bfloat16)
            mk = ops.Mask(y1, *(torch.rand(Cout, generator=g, device=dev) + 0.5 for _ in range(4)))
        tail = None
        if kind == "fwd":
            f32 = lambda: torch.zeros(Cout, device=dev)
            counter = torch.zeros(1, dtype=torch.int32, device=dev)
            tail = ops.bn_tail(counter, float(N * H * W), f32() + 1, f32(), f32(), f32(), f32() + 1, 0.1, 1e-5, f32(), f32(), f32(), f32())
        e0.record(); ops.conv3x3(N, H, W, sources, Cout, wp, want_stats=(kind == "fwd"), mask=mk, tail=tail); e1.record()
        torch.cuda.synchronize(); rd(buf, 0)
    v = [x / (plan.grid_m * plan.n_tiles) / 1e3 for x in buf]   # kcycles per CTA
    ms = e0.elapsed_time(e1)
    raw = list(buf)
    fl = 2.0 * N * H * W * cin * Cout * 9
    print("%s %dx%d cin %d cout %d: %.3f ms %.0f TF/s | T=%d NT=%d CK=%d S=%d chunks=%d groups/cta=%.1f" % (kind, H, W, cin, Cout, ms, fl / ms / 1e9, plan.T, plan.NT, plan.CK, plan.S, plan.nchunks, plan.ntg / plan.grid_m))
    print("   kcycles/CTA  producer: wait_acc %.0f table %.0f wait_empty %.0f stage %.0f | epilogue: wait %.0f work %.0f | mma: wait_acc %.0f wait_full %.0f issue %.0f | fetch: wait_empty %.0f issue %.0f"
          % (v[0], v[1], v[2], v[3], v[4], v[5], v[8], v[9], v[10], v[12], v[13]))
    print("   mma prologue (entry -> first wait) %.1f kcycles/CTA | epilogue after the last group, slowest CTA (statistics + BatchNorm tail) %.1f kcycles" % (v[11], raw[6] / 1e3))
