"""Host-side enqueue time per training step vs GPU time (is the step launch-bound?)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import archs, losses, ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
model = archs.NestedUNet(1, 3, True).cuda()
crit = losses.BCEDiceLoss()
opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
x = torch.randn(B, 3, 96, 96, device="cuda"); t = (torch.rand(B, 1, 96, 96, device="cuda") > 0.5).float()
def step():
    out = model(x); loss = sum(crit(o, t) for o in out) / 4
    opt.zero_grad(); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
for B_ in (B,):
    t0 = time.perf_counter()
    for _ in range(5): step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("batch %d: host enqueue %.2f ms/step, total %.2f ms/step" % (B_, (t1 - t0) / 5 * 1e3, (t2 - t0) / 5 * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
