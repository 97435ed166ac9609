"""Times a few level-0/1 conv launches of the generic kernel (experiment builds: tools/ab_epilogue.sh)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import ops, _lib as L
dev = "cuda:0"; N = 256
g = torch.Generator(device=dev).manual_seed(0)
def timeit(fn):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 8
res = []
for H, cin, cout, kind in [(48, 64, 320, "plain"), (48, 64, 256, "plain"), (48, 64, 192, "plain"), (24, 128, 512, "plain"), (24, 128, 384, "plain"), (24, 128, 256, "plain"), (12, 256, 768, "plain"), (12, 256, 512, "plain"), (12, 256, 256, "stats"), (12, 256, 256, "mask"), (12, 768, 256, "stats"), (12, 128, 256, "stats"), (6, 512, 512, "stats"), (6, 256, 512, "stats"), (12, 256, 768, "plain"), (6, 512, 512, "mask"), (96, 192, 32, "c32stats"), (96, 96, 32, "c32stats"), (96, 192, 32, "c32plain"), (96, 96, 32, "c32plain"), (96, 64, 32, "c32plain"),(96, 32, 32, "stats"), (96, 32, 32, "plain"), (96, 32, 32, "mask"), (96, 32, 96, "plain"), (96, 32, 192, "plain"), (96, 32, 192, "segs"), (48, 64, 64, "stats"), (48, 64, 320, "plain")]:
    a = torch.rand(N, H, H, cin, generator=g, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, generator=g, device=dev) * 0.05
    plan = ops.conv_plan(N, H, H, cin, cout)
    wp = ops.pack_weights(w, L.PACK_FWD, plan)
    src = [ops.Source(L.SRC_ACT, a, cin)]
    if kind.startswith("c32"):
        wpc = ops.pack_weights(w, L.PACK_FWD_C32, plan)
        if kind == "c32stats": fn = lambda: ops.conv3x3(N, H, H, src, cout, wpc, want_stats=True, stacked=True)
        elif kind == "c32plain": fn = lambda: ops.conv3x3(N, H, H, src, cout, wpc, stacked=True)
        else:
            y1 = torch.randn(N, H, H, cout, generator=g, device=dev).to(torch.bfloat16)
            sc, sh, mu, rs = (torch.rand(cout, generator=g, device=dev) + 0.5 for _ in range(4))
            fn = lambda: ops.conv3x3(N, H, H, src, cout, wpc, mask=ops.Mask(y1, sc, sh, mu, rs), stacked=True)
    elif kind == "stats": fn = lambda: ops.conv3x3(N, H, H, src, cout, wp, want_stats=True)
    elif kind == "plain": fn = lambda: ops.conv3x3(N, H, H, src, cout, wp)
    elif kind == "segs": fn = lambda: ops.conv3x3(N, H, H, src, cout, wp, out_segs=[32] * 4 + [64])
    else:
        y1 = torch.randn(N, H, H, cout, generator=g, device=dev).to(torch.bfloat16)
        sc, sh, mu, rs = (torch.rand(cout, generator=g, device=dev) + 0.5 for _ in range(4))
        fn = lambda: ops.conv3x3(N, H, H, src, cout, wp, mask=ops.Mask(y1, sc, sh, mu, rs))
    res.append("%d %d->%d %s %.3f" % (H, cin, cout, kind, timeit(fn)))
print(os.environ.get("TAG", ""), " | ".join(res), flush=True)
