"""Times (or, under ncu with ONCE=1, runs once) the HBM-bound kernels of the step on their level-0/1 shapes at batch 256."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import ops, _lib as L
dev = "cuda:0"
N = int(os.environ.get("N", "256"))
ONCE = os.environ.get("ONCE", "0") == "1"
g = torch.Generator(device=dev).manual_seed(0)
def bf(*shape): return torch.randn(*shape, generator=g, device=dev).to(torch.bfloat16)
def f32(*shape): return torch.randn(*shape, generator=g, device=dev)
def run(name, nbytes, fn):
    reps = 1 if ONCE else 10
    if not ONCE:
        for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-44s %.3f ms  %6.0f GB/s (algorithmic %.0f MB)" % (name, ms, nbytes / ms / 1e6, nbytes / 1e6))
H = 96; C = 32
y = bf(N, H, H, C); sc, sh, mu, rs = (f32(C).abs() + 0.5 for _ in range(4))
U = N * H * H * C * 2
# router: 1 dense consumer (conv1 output <- conv2 dgrad)
dx1 = bf(N, H, H, 32)
run("router 96 C32 1 dense consumer", 3 * U, lambda: ops.grad_router(N, H, H, C, y, sc, sh, mu, rs, [ops.Consumer(dx1, 32, 0, L.SRC_SAME)]))
# router: 3 same consumers (pitches 96,128,160... offsets 32) + head
dxs = [bf(N, H, H, p) for p in (128, 160, 192)]
dl = f32(N, 1, H, H); wh = f32(1, C)
run("router 96 C32 3 same + head", 5 * U + N * H * H * 4, lambda: ops.grad_router(N, H, H, C, y, sc, sh, mu, rs, [ops.Consumer(d, d.shape[-1], 32, L.SRC_SAME) for d in dxs], dlogits=dl, wh=wh))
dxs4 = [bf(N, H, H, p) for p in (96, 128, 160, 192)]
dxp = bf(N, H // 2, H // 2, 32)
run("router 96 C32 4 same + pool", 6 * U + U // 4, lambda: ops.grad_router(N, H, H, C, y, sc, sh, mu, rs, [ops.Consumer(d, d.shape[-1], 0, L.SRC_SAME) for d in dxs4] + [ops.Consumer(dxp, 32, 0, L.SRC_POOL)]))
# level 1 router with 2 same consumers
y1 = bf(N, 48, 48, 64); sc1, sh1, mu1, rs1 = (f32(64).abs() + 0.5 for _ in range(4))
d1 = [bf(N, 48, 48, p) for p in (256, 320)]
run("router 48 C64 2 same", 4 * U // 2, lambda: ops.grad_router(N, 48, 48, 64, y1, sc1, sh1, mu1, rs1, [ops.Consumer(d, d.shape[-1], 64, L.SRC_SAME) for d in d1]))
# up_transpose: level-0 dX (pitch 192, slice of 64 ch at offset 128) -> level 1
dxu = bf(N, H, H, 192)
run("up_transpose 96->48 C64 pitch 192", 2 * U + U // 2, lambda: ops.up_transpose(dxu, 192, 128, N, 48, 48, 64))
a1 = bf(N, 48, 48, 64)
run("upsample2x 48->96 C64", 2 * U + U // 2, lambda: ops.upsample2x(a1))
run("bn_apply 96 C32", 2 * U, lambda: ops.bn_apply(y, sc, sh))
run("head_bwd_params 96 C32", U + N * H * H * 4, lambda: ops.head_bwd_params(y, sc, sh, dl, N, H, H, 1))
x = f32(N, 3, H, H)
run("input_to_nhwc32", U + N * 3 * H * H * 4, lambda: ops.input_to_nhwc32(x))
a0 = bf(N, H, H, C)
run("maxpool2x2 96 C32", U + U // 4, lambda: ops.maxpool2x2(a0))
