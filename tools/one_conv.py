import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import ops, _lib as L
dev = "cuda:0"
N, H, cin, cout = int(os.environ.get("N", "8")), int(os.environ.get("H", "24")), int(os.environ.get("CIN", "128")), int(os.environ.get("COUT", "128"))
g = torch.Generator(device=dev).manual_seed(0)
a = torch.rand(N, H, H, cin, generator=g, device=dev).to(torch.bfloat16)
w = torch.randn(cout, cin, 3, 3, generator=g, device=dev) * 0.05
plan = ops.conv_plan(N, H, H, cin, cout)
print("plan T", plan.T, "NT", plan.NT, "CK", plan.CK, "S", plan.S, "tma", plan.tma_smem_bytes, "epi", plan.tma_epi_warps, flush=True)
stacked = os.environ.get("STACKED") == "1"
wp = ops.pack_weights(w, L.PACK_FWD_C32 if stacked else L.PACK_FWD, plan)
src = [ops.Source(L.SRC_ACT, a, cin)]
for kind in os.environ.get("KINDS", "plain,stats").split(","):
    y, part = ops.conv3x3(N, H, H, src, cout, wp, want_stats=(kind == "stats"), stacked=stacked)
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(a.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
    print(kind, "ok maxerr", float((y.float() - ref).abs().max()), flush=True)
