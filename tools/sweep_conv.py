"""Sweeps the generic conv kernel's tiling (T tiles per group, S stages, resident weight block) on the step's layer shapes.
Needs a library built with -DNUNET_EXPERIMENT (tools/sweep_conv.sh builds it on the GPU box)."""
import itertools, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import ops, _lib as L
dev = "cuda:0"
N = int(os.environ.get("N", "256"))
g = torch.Generator(device=dev).manual_seed(0)
# (H, cin, cout, stats?)
SH = [(96, 32, 32, True), (96, 32, 32, False), (96, 32, 96, False), (96, 32, 128, False), (96, 32, 160, False), (96, 32, 192, False),
      (48, 32, 64, True), (48, 64, 64, True), (48, 64, 64, False), (48, 64, 192, False), (48, 64, 320, False), (48, 192, 64, True), (48, 320, 64, True),
      (24, 64, 128, True), (24, 128, 128, True), (24, 128, 384, False), (24, 128, 512, False), (24, 512, 128, True),
      (12, 128, 256, True), (12, 256, 256, True), (12, 256, 768, False), (12, 768, 256, True), (6, 256, 512, True), (6, 512, 512, True), (6, 512, 512, False)]
def timeit(fn):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 8
for H, cin, cout, stats in SH:
    a = torch.rand(N, H, H, cin, generator=g, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, generator=g, device=dev) * 0.05
    res = []
    for T, S, rb in itertools.product(("", "1", "2", "4"), ("", "2", "3", "4"), ("", "32")):
        if (T == "") != (S == ""): continue
        for k, v in (("NUNET_CONV_T", T), ("NUNET_CONV_S", S), ("NUNET_RESB_MIN_NT", rb)):
            if v: os.environ[k] = v
            else: os.environ.pop(k, None)
        ops.conv_plan.cache_clear()
        try:
            plan = ops.conv_plan(N, H, H, cin, cout)
            if T and (plan.T != int(T)): continue
            wp = ops.pack_weights(w, L.PACK_FWD, plan)
            ms = timeit(lambda: ops.conv3x3(N, H, H, [ops.Source(L.SRC_ACT, a, cin)], cout, wp, want_stats=stats))
        except Exception as e:
            continue
        res.append((ms, "T=%d S=%d CK=%d nch=%d resb=%s%s" % (plan.T, plan.S, plan.CK, plan.nchunks, rb or "128", "" if T else " (default)")))
    fl = 2.0 * N * H * H * cin * cout * 9
    by = 2.0 * N * H * H * (cin + cout)
    base = [r for r in res if "default" in r[1] and "resb=128" in r[1]][0][0]
    res.sort()
    print("conv %3d %4d->%4d %s default %.3f ms | roofline tc %.3f hbm %.3f | best: %s" % (H, cin, cout, "stats" if stats else "plain", base, fl / 1390.8e9, by / 6541e6,
          "  ".join("%.3f %s" % r for r in res[:4])), flush=True)
