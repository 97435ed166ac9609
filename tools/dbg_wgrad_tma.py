import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ops, _lib as L
dev = "cuda:0"
def bf(t): return t.to(torch.bfloat16).float()
CASES = [(2, 32, 32, [32, 32, 32], 32), (8, 96, 96, [32, 64], 32), (3, 48, 80, [32] * 5, 32), (2, 16, 16, [64, 64, 128], 64),
         (2, 24, 24, [128, 256], 128), (3, 12, 20, [256, 512], 256), (2, 6, 6, [512], 512), (5, 6, 6, [256], 512), (1, 16, 16, [32], 32)]
for N, H, W, Cs, Cout in CASES:
    g = torch.Generator(device=dev).manual_seed(1)
    srcs, xs = [], []
    for Cc in Cs:
        a = torch.rand(N, H, W, Cc, generator=g, device=dev).to(torch.bfloat16)
        srcs.append(ops.Source(L.SRC_ACT, a, Cc)); xs.append(a.float().permute(0, 3, 1, 2))
    cin = sum(Cs)
    gy = (torch.randn(N, H, W, Cout, generator=g, device=dev)).to(torch.bfloat16)
    dw = torch.full((Cout, cin, 3, 3), float("nan"), device=dev)
    pl = ops.wgrad_plan(N, H, W, cin, Cout)
    ops.conv3x3_wgrad(N, H, W, srcs, Cout, gy, dw)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(torch.cat(xs, 1), (Cout, cin, 3, 3), gy.float().permute(0, 3, 1, 2), padding=1)
    err = (dw - ref).abs()
    nan = int(torch.isnan(dw).sum())
    e2 = torch.nan_to_num(err, nan=0.0)
    # per-tap / per-ci-block / per-co-block error maps
    taperr = e2.amax((0, 1)).flatten().tolist()
    print("N%d %dx%d cin %s cout %d | tma=%d PW=%d r=%d S=%d G=%d Nt=%d splits=%d | nan %d maxerr %.3g (ref max %.3g)" % (
        N, H, W, Cs, Cout, pl.tma, pl.tma_PW, pl.tma_r, pl.tma_S, pl.tma_G, pl.tma_Nt, pl.tma_splits, nan, float(e2.max()), float(ref.abs().max())))
    print("    tap errs", ["%.2g" % v for v in taperr], " ci-block errs", ["%.2g" % float(e2[:, i:i + 32].max()) for i in range(0, cin, 32)][:8],
          " co-block errs", ["%.2g" % float(e2[i:i + 32].max()) for i in range(0, Cout, 32)][:8])
