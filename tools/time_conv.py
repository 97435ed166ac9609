"""Times forward / dgrad conv launches (copy-mode sources) on the level-0/1 shapes at batch 256."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import ops, _lib as L
dev = "cuda:0"
N = int(os.environ.get("N", "256"))
g = torch.Generator(device=dev).manual_seed(0)
SH = [(96, 32, 32), (96, 32, 96), (96, 32, 128), (96, 32, 160), (96, 32, 192), (96, 96, 32), (96, 192, 32), (48, 64, 64), (48, 64, 320), (48, 320, 64), (24, 128, 512)]
for H, cin, cout in SH:
    a = torch.rand(N, H, H, cin, generator=g, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, generator=g, device=dev) * 0.05
    plan = ops.conv_plan(N, H, H, cin, cout)
    wp = ops.pack_weights(w, L.PACK_FWD, plan)
    for stats in (False, True):
        fn = lambda: ops.conv3x3(N, H, H, [ops.Source(L.SRC_ACT, a, cin)], cout, wp, want_stats=stats)
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * N * H * H * cin * cout * 9
        print("conv %3d cin %4d cout %4d %s: %.3f ms %5.0f TF/s | T=%d NT=%d S=%d nch=%d" % (H, cin, cout, "fwd+stats" if stats else "plain    ", ms, fl / ms / 1e9, plan.T, plan.NT, plan.S, plan.nchunks))
