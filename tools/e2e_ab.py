"""A/B of end-to-end loop variants (host->device copy placement) on the headline config."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import archs, losses, ops, utils
dev = torch.device("cuda:0")
B = 256
torch.manual_seed(0)
model = archs.NestedUNet(1, 3, True).to(dev).train()
crit = losses.BCEDiceLoss()
opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
xh = (torch.randn(B, 3, 96, 96) * 0.01).pin_memory(); th = (torch.rand(B, 1, 96, 96) > 0.5).float().pin_memory()
def step(x, t):
    out = model(x); loss = sum(crit(o, t) for o in out) / 4
    c = ops.iou_counts(out[-1].detach(), t)
    opt.zero_grad(); loss.backward(); opt.step()
    return loss, c
xd, td = xh.to(dev), th.to(dev)
for _ in range(5): step(xd, td)
torch.cuda.synchronize()
K = 20
def timeit(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    print("%-40s %.2f ms/step" % (name, (time.perf_counter() - t0) / K * 1e3))
def resident():
    for _ in range(K): step(xd, td)
def resident_sync():
    for _ in range(K):
        l, c = step(xd, td); l.item(); c.tolist()
def old():
    for _ in range(K):
        x = xh.to(dev, non_blocking=True); t = th.to(dev, non_blocking=True)
        l, c = step(x, t); l.item(); c.tolist()
def pref():
    def gen():
        for _ in range(K): yield xh, th, None
    for x, t, _ in utils.DevicePrefetcher(gen(), dev):
        l, c = step(x, t); l.item(); c.tolist()
cs = torch.cuda.Stream()
def late():
    # copy of the next batch issued after the step is enqueued
    with torch.cuda.stream(cs):
        nx, nt = xh.to(dev, non_blocking=True), th.to(dev, non_blocking=True)
    for _ in range(K):
        torch.cuda.current_stream().wait_stream(cs)
        x, t = nx, nt
        l, c = step(x, t)
        with torch.cuda.stream(cs):
            nx, nt = xh.to(dev, non_blocking=True), th.to(dev, non_blocking=True)
        l.item(); c.tolist()
for name, fn in (("resident", resident), ("resident + host sync", resident_sync), ("copy on the compute stream", old), ("prefetcher", pref), ("copy after enqueue", late),
                 ("resident", resident), ("prefetcher", pref)):
    timeit(name, fn)
