"""Runs ONE launch of each kernel class of the step on its level-0/1 shape at batch 256 (for `ncu --set full --import-source on`)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import ops, _lib as L
dev = "cuda:0"
N = int(os.environ.get("N", "256"))
WHICH = os.environ.get("WHICH", "all").split(",")
g = torch.Generator(device=dev).manual_seed(0)
def bf(*shape): return torch.randn(*shape, generator=g, device=dev).to(torch.bfloat16)
def f32(*shape): return torch.randn(*shape, generator=g, device=dev)
def want(k): return "all" in WHICH or k in WHICH
def conv(H, cin, cout, stats, mode=L.PACK_FWD, stacked=False):
    a = bf(N, H, H, cin).abs()
    w = f32(cout, cin, 3, 3) * 0.05
    plan = ops.conv_plan(N, H, H, cin, cout)
    wp = ops.pack_weights(w, mode, plan)
    ops.conv3x3(N, H, H, [ops.Source(L.SRC_ACT, a, cin)], cout, wp, want_stats=stats, stacked=stacked)
if want("dgrad192"): conv(96, 32, 192, False)
if want("fwd32"): conv(96, 32, 32, True)
if want("c32"): conv(96, 192, 32, True, L.PACK_FWD_C32, True)
if want("dgrad320"): conv(48, 64, 320, False)
H = 96; C = 32
y = bf(N, H, H, C); sc, sh, mu, rs = (f32(C).abs() + 0.5 for _ in range(4))
if want("router"):
    dxs4 = [bf(N, H, H, p) for p in (96, 128, 160, 192)]
    dxp = bf(N, H // 2, H // 2, 32)
    ops.grad_router(N, H, H, C, y, sc, sh, mu, rs, [ops.Consumer(d, d.shape[-1], 0, L.SRC_SAME) for d in dxs4] + [ops.Consumer(dxp, 32, 0, L.SRC_POOL)])
    dl = f32(N, 1, H, H); wh = f32(1, C)
    ops.grad_router(N, H, H, C, y, sc, sh, mu, rs, [ops.Consumer(d, d.shape[-1], 32, L.SRC_SAME) for d in dxs4[1:]], dlogits=dl, wh=wh)
    del dxs4, dxp
if want("up"):
    dxu = bf(N, H, H, 192)
    ops.up_transpose(dxu, 192, 128, N, 48, 48, 64)
    a1 = bf(N, 48, 48, 64)
    ops.upsample2x(a1)
    del dxu, a1
if want("wgrad"):
    for HH, Cs, Cout in ((96, [32] * 4 + [64], 32), (48, [64], 64), (96, [32], 32)):
        srcs = [ops.Source(L.SRC_ACT, bf(N, HH, HH, c), c) for c in Cs]
        gy = bf(N, HH, HH, Cout)
        dw = torch.empty(Cout, sum(Cs), 3, 3, device=dev)
        ops.conv3x3_wgrad(N, HH, HH, srcs, Cout, gy, dw)
torch.cuda.synchronize()
print("ok")
