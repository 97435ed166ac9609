import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import archs, engine
torch.manual_seed(0)
model = archs.NestedUNet(3, 3, True).cuda()
gen = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(4, 3, 96, 96, generator=gen, device="cuda")
with torch.no_grad():
    _, recs, _ = engine._forward(model, x, True)
specs = engine._state(model).specs
for s in specs:
    r = recs[s.idx]
    y = r.y.float().permute(0, 3, 1, 2)
    ze = y * r.scale.view(1, -1, 1, 1) + r.shift.view(1, -1, 1, 1)
    zt = F.batch_norm(y, None, None, s.bn.weight, s.bn.bias, True, 0.0, s.bn.eps)
    m = y.mean((0, 2, 3)); v = y.var((0, 2, 3), unbiased=False)
    flips = int(((ze > 0) != (zt > 0)).sum())
    print("%-6s w%d  max|ze-zt| %.3e  flips %d / %d  mean err %.2e  invstd relerr %.2e" % (
        s.node, s.which, float((ze - zt).abs().max()), flips, ze.numel(),
        float((m - r.mean).abs().max()), float(((1 / torch.sqrt(v + 1e-5)) / r.invstd - 1).abs().max())))
