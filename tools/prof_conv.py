"""Runs a single fused conv / wgrad / router launch configuration a few times (target for ncu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import ops, _lib as L
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
N, H, W, Cin, Cout = [int(v) for v in (sys.argv[2:7] if len(sys.argv) > 6 else (256, 96, 96, 32, 32))]
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
y = torch.randn(N, H, W, Cin, generator=g, device=dev).to(torch.bfloat16)
sc = torch.randn(Cin, generator=g, device=dev); sh = torch.randn(Cin, generator=g, device=dev)
w = torch.randn(Cout, Cin, 3, 3, generator=g, device=dev) * 0.05
src = ops.Source(L.SRC_ACT if os.environ.get('ACT') else L.SRC_SAME, y, Cin, sc, sh)
gz = torch.randn(N, H, W, Cout, generator=g, device=dev).to(torch.bfloat16)
yo = torch.randn(N, H, W, Cout, generator=g, device=dev).to(torch.bfloat16)
k1, k2, k3 = (torch.randn(Cout, generator=g, device=dev) for _ in range(3))
mean = torch.randn(Cout, generator=g, device=dev); inv = torch.rand(Cout, generator=g, device=dev) + 0.5
for it in range(3):
    if which == "fwd":
        plan = ops.conv_plan(N, H, W, Cin, Cout)
        wp = ops.pack_weights(w, L.PACK_FWD, plan)
        ops.conv3x3(N, H, W, [src], Cout, wp, want_stats=True)
    elif which == "wgrad":
        dw = torch.empty(Cout, Cin, 3, 3, device=dev)
        ops.conv3x3_wgrad(N, H, W, [src], Cout, ops.bn_bwd_apply(gz, yo, k1, k2, k3), dw)
    elif which == "router":
        dx = torch.randn(N, H, W, Cout, generator=g, device=dev).to(torch.bfloat16)
        ops.grad_router(N, H, W, Cout, yo, k1, k2, mean, inv, [ops.Consumer(dx, Cout, 0, L.SRC_SAME)])
    torch.cuda.synchronize()
print("done", which)
