"""Parity + timing of the stacked Cout=32 conv kernel (conv_c32.cu) against torch on identical bf16 operands."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import ops, _lib as L
dev = "cuda:0"
CASES = [(2, 16, 16, [32]), (2, 32, 32, [32, 64]), (3, 48, 80, [32, 32, 64]), (4, 96, 96, [32] * 4 + [64]), (2, 6, 6, [64]), (1, 16, 16, [32])]
TIME = os.environ.get("TIME", "0") == "1"
if TIME:
    CASES = [(256, 96, 96, [32, 64]), (256, 96, 96, [32, 32, 64]), (256, 96, 96, [32] * 3 + [64]), (256, 96, 96, [32] * 4 + [64])]
for N, H, W, Cs in CASES:
    g = torch.Generator(device=dev).manual_seed(1)
    cin = sum(Cs)
    srcs, xs = [], []
    for c in Cs:
        a = (torch.rand(N, H, W, c, generator=g, device=dev) - 0.3).to(torch.bfloat16)
        srcs.append(ops.Source(L.SRC_ACT, a, c)); xs.append(a.float().permute(0, 3, 1, 2))
    if not ops.c32_eligible(N, H, W, cin, 32):
        print('not eligible:', N, H, W, cin)
        continue
    plan = ops.conv_plan(N, H, W, cin, 32)
    for mode in ("fwd", "dgrad"):
        if mode == "fwd":
            w = torch.randn(32, cin, 3, 3, generator=g, device=dev) * 0.1
            wp = ops.pack_weights(w, L.PACK_FWD_C32, plan)
            wp0 = ops.pack_weights(w, L.PACK_FWD, plan)
            ref = F.conv2d(torch.cat(xs, 1), w.to(torch.bfloat16).float(), padding=1)
        else:
            w = torch.randn(cin, 32, 3, 3, generator=g, device=dev) * 0.1     # [K = Cout of the forward conv, N = 32 input channels]
            wp = ops.pack_weights(w, L.PACK_DGRAD_C32, plan)
            wp0 = ops.pack_weights(w, L.PACK_DGRAD, plan)
            ref = F.conv_transpose2d(torch.cat(xs, 1), w.to(torch.bfloat16).float(), padding=1)
        y, part = ops.conv3x3(N, H, W, srcs, 32, wp, want_stats=True, stacked=True)
        y0, part0 = ops.conv3x3(N, H, W, srcs, 32, wp0, want_stats=True)
        torch.cuda.synchronize()
        refn = ref.permute(0, 2, 3, 1)
        err = float((y.float() - refn).abs().max()); err0 = float((y0.float() - refn).abs().max())
        s_ref = refn.reshape(-1, 32).double().sum(0); q_ref = (refn.reshape(-1, 32).double() ** 2).sum(0)
        s = part[:, 0].double().sum(0); q = part[:, 1].double().sum(0)
        print("%s N%d %dx%d cin %d: max err %.4g (generic kernel %.4g, ref max %.3g) | stats rel err sum %.2e sq %.2e | nan %d" % (
            mode, N, H, W, cin, err, err0, float(refn.abs().max()), float(((s - s_ref).abs() / (s_ref.abs() + 1)).max()),
            float(((q - q_ref).abs() / (q_ref.abs() + 1)).max()), int(torch.isnan(y.float()).sum())))
        if TIME:
            for st in (True, False):
                ws = os.environ.get('NUNET_C32_GRID') is None
                for _ in range(3): ops.conv3x3(N, H, W, srcs, 32, wp if st else wp0, want_stats=ws, stacked=st)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10): ops.conv3x3(N, H, W, srcs, 32, wp if st else wp0, want_stats=ws, stacked=st)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                print("    %s: %.3f ms  %.0f TF/s" % ("stacked" if st else "generic", ms, 2.0 * N * H * W * cin * 32 * 9 / ms / 1e9))
