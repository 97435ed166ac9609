import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import archs, engine
torch.manual_seed(0)
model = archs.NestedUNet(3, 3, True).cuda()
gen = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(4, 3, 96, 96, generator=gen, device="cuda")
specs = engine._state(model).specs
with torch.no_grad():
    _, r1, _ = engine._forward(model, x, True)
    torch.cuda.synchronize()
    for trial in range(3):
        _, r2, _ = engine._forward(model, x, True)
        torch.cuda.synchronize()
        for s in specs:
            a, b = r1[s.idx].y, r2[s.idx].y
            ne = (a != b)
            if bool(ne.any()):
                idx = ne.nonzero()
                print("trial %d layer %s w%d: %d / %d differ; max abs diff %.3e; first idx %s; distinct n %s, distinct c count %d, y range %s x range %s" % (
                    trial, s.node, s.which, int(ne.sum()), a.numel(), float((a.float() - b.float()).abs().max()), idx[0].tolist(),
                    sorted(set(idx[:, 0].tolist())), len(set(idx[:, 3].tolist())), (int(idx[:, 1].min()), int(idx[:, 1].max())), (int(idx[:, 2].min()), int(idx[:, 2].max()))))
                sc = (r1[s.idx].scale != r2[s.idx].scale).sum()
                print("    scale mismatches:", int(sc))
                break
        else:
            print("trial %d: all layers bitwise identical" % trial)
