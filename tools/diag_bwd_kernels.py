"""Runs one engine step and re-checks every wgrad / dgrad launch against torch on the SAME real operands."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import archs, losses, ops, _lib as L  # noqa: E402
from refops import source_tensor  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
KIND = {L.SRC_SAME: "same", L.SRC_POOL: "pool", L.SRC_UP: "up"}


def bf(t):
    return t.to(torch.bfloat16).float()


orig_wgrad, orig_conv = ops.conv3x3_wgrad, ops.conv3x3


def gy_of(gz, y, k1, k2, k3):
    return (gz.float() * k1 + y.float() * k2 + k3).permute(0, 3, 1, 2)


def wgrad_checked(N, H, W, sources, Cout, gz, y, k1, k2, k3, dw, part=None):
    orig_wgrad(N, H, W, sources, Cout, gz, y, k1, k2, k3, dw, part)
    xs = []
    for s in sources:
        if s.kind == L.SRC_INPUT:
            xs.append(s.data)
        else:
            xs.append(source_tensor(KIND[s.kind], s.data, s.scale, s.shift))
    x = torch.cat(xs, 1)
    gy = gy_of(gz, y, k1, k2, k3)
    ref_r = torch.nn.grad.conv2d_weight(bf(x), dw.shape, bf(gy), padding=1)
    ref_x = torch.nn.grad.conv2d_weight(x, dw.shape, gy, padding=1)
    e1 = float((dw - ref_r).norm() / ref_r.norm())
    e2 = float((ref_r - ref_x).norm() / ref_x.norm())
    print("wgrad %dx%d cin %d cout %d: kernel vs torch(bf16 operands) %.3e | bf16-operand rounding effect %.3e | sum(gy)/sum|gy| %.2e"
          % (H, W, x.shape[1], Cout, e1, e2, float(gy.sum().abs() / gy.abs().sum())))
    return dw


def conv_checked(N, H, W, sources, Cout, wpack, out=None, stat_part=None, want_stats=False):
    res = orig_conv(N, H, W, sources, Cout, wpack, out, stat_part, want_stats)
    if sources[0].kind == L.SRC_GRAD:
        conv_checked.pending = (N, H, W, sources[0], Cout, res[0])
    return res


ops.conv3x3_wgrad = wgrad_checked
import engine  # noqa: E402
engine.ops.conv3x3_wgrad = wgrad_checked

torch.manual_seed(0)
model = archs.NestedUNet(3, 3, True).cuda()
gen = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(4, 3, 96, 96, generator=gen, device="cuda")
t = (torch.rand(4, 3, 96, 96, generator=gen, device="cuda") > 0.5).float()
crit = losses.BCEDiceLoss()
out = model(x)
loss = sum(crit(o, t) for o in out) / 4
loss.backward()
