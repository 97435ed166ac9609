"""Times the weight-gradient kernel on the NestedUNet layer shapes (all sources pre-activated) at batch 256."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import ops, _lib as L
dev = "cuda:0"
N = int(os.environ.get("N", "256"))
SH = [(96, [32], 32), (96, [32, 64], 32), (96, [32] * 4 + [64], 32), (48, [32], 64), (48, [64], 64), (48, [64, 128], 64), (48, [64] * 3 + [128], 64),
      (24, [64], 128), (24, [128], 128), (24, [128, 256], 128), (24, [128] * 2 + [256], 128), (12, [128], 256), (12, [256], 256),
      (12, [256, 512], 256), (6, [256], 512), (6, [512], 512)]
g = torch.Generator(device=dev).manual_seed(0)
tot = 0.0
for H, Cs, Cout in SH:
    srcs = [ops.Source(L.SRC_ACT, torch.rand(N, H, H, c, generator=g, device=dev).to(torch.bfloat16), c) for c in Cs]
    cin = sum(Cs)
    gy = torch.randn(N, H, H, Cout, generator=g, device=dev).to(torch.bfloat16)
    dw = torch.empty(Cout, cin, 3, 3, device=dev)
    pl = ops.wgrad_plan(N, H, H, cin, Cout)
    for _ in range(3):
        ops.conv3x3_wgrad(N, H, H, srcs, Cout, gy, dw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.conv3x3_wgrad(N, H, H, srcs, Cout, gy, dw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * N * H * H * cin * Cout * 9
    by = 2.0 * N * H * H * (cin + Cout)
    tot += ms
    print("wgrad %3dx%-3d cin %4d cout %4d: %.3f ms %5.0f TF/s %5.0f GB/s | tma=%d PW=%d r=%d S=%d G=%d splits=%d cpc=%d" % (
        H, H, cin, Cout, ms, fl / ms / 1e9, by / ms / 1e6, pl.tma, pl.tma_PW, pl.tma_r, pl.tma_S, pl.tma_G, pl.tma_splits, pl.tma_cpc))
print("sum %.3f ms" % tot)
