"""One weight-gradient launch at a BASELINE shape (for ncu) + timing."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import ops, _lib as L
dev = "cuda:0"
N, H, cin, cout = int(os.environ.get("N", "256")), int(os.environ.get("H", "96")), int(os.environ.get("CIN", "32")), int(os.environ.get("COUT", "32"))
g = torch.Generator(device=dev).manual_seed(0)
a = torch.rand(N, H, H, cin, generator=g, device=dev).to(torch.bfloat16)
gy = torch.randn(N, H, H, cout, generator=g, device=dev).to(torch.bfloat16)
dw = torch.empty(cout, cin, 3, 3, device=dev)
src = [ops.Source(L.SRC_ACT, a, cin)]
pl = ops.wgrad_plan(N, H, H, cin, cout)
print("plan tma", pl.tma, "PW", pl.tma_PW, "r", pl.tma_r, "S", pl.tma_S, "G", pl.tma_G, "splits", pl.tma_splits, "cpc", pl.tma_cpc, "chunks", pl.tma_chunks, "smem", pl.tma_smem_bytes, flush=True)
reps = int(os.environ.get("REPS", "8"))
for _ in range(2): ops.conv3x3_wgrad(N, H, H, src, cout, gy, dw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): ops.conv3x3_wgrad(N, H, H, src, cout, gy, dw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("wgrad %dx%d %d->%d: %.3f ms (incl. split-K reduce), %.0f GB/s algorithmic, %.0f TFLOP/s" % (H, H, cin, cout, ms, (a.numel() + gy.numel()) * 2 / ms / 1e6, 2.0 * N * H * H * cin * cout * 9 / ms / 1e9), flush=True)
