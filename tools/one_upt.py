"""Times the transposed-upsample kernel at the shapes of the BASELINE step (one launch each for ncu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import ops
dev = "cuda:0"; N = 256
g = torch.Generator(device=dev).manual_seed(0)
reps = int(os.environ.get("REPS", "8"))
for h, C in [(48, 64), (24, 128), (12, 256), (6, 512)]:
    dx = torch.randn(N, 2 * h, 2 * h, C, generator=g, device=dev).to(torch.bfloat16)
    out = torch.empty(N, h, h, C, dtype=torch.bfloat16, device=dev)
    for _ in range(2): ops.up_transpose(dx, C, 0, N, h, h, C, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.up_transpose(dx, C, 0, N, h, h, C, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    by = (dx.numel() + out.numel()) * 2
    print("up_transpose %dx%d C=%d: %.3f ms, %.0f GB/s algorithmic" % (h, h, C, ms, by / ms / 1e6), flush=True)
