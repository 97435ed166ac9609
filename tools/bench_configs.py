"""Throughput of the other BASELINE.json configurations (C1, C3, C4, C5) on one GPU; writes gpurun_out/configs.json (copied to
profiles/rNN_configs.json).  fp32 NCHW inputs as the reference loop feeds them; the package's fused SGD."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import archs, losses, metrics, ops
import optim as noptim

dev = torch.device("cuda:0")


def train_rate(arch, classes, ds, n, size, steps=5, warm=3):   # warm >= 3: the second step captures the CUDA graphs
    torch.manual_seed(0)
    model = getattr(archs, arch)(classes, 3, ds).to(dev)
    crit = losses.BCEDiceLoss()
    opt = noptim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    x = torch.randn(n, 3, size, size, device=dev) * 0.01
    t = (torch.rand(n, classes, size, size, device=dev) > 0.5).float()

    def step():
        out = model(x)
        outs = out if isinstance(out, list) else [out]
        loss = sum(crit(o, t) for o in outs) / len(outs)
        iou = metrics.iou_counts(outs[-1], t)
        opt.zero_grad(); loss.backward(); opt.step()
        return loss
    for _ in range(warm): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): loss = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model, opt
    torch.cuda.empty_cache()
    return {"images_per_s": n / ms * 1e3, "ms_per_step": ms, "loss": float(loss)}


def eval_rate(arch, n, size, steps=10, warm=3):
    torch.manual_seed(0)
    model = getattr(archs, arch)(1, 3, False).to(dev).eval()
    x = torch.randn(n, 3, size, size, device=dev) * 0.01
    t = (torch.rand(n, 1, size, size, device=dev) > 0.5).float()
    with torch.no_grad():
        for _ in range(warm):
            metrics.iou_score(model(x), t)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            iou = metrics.iou_score(model(x), t)      # val.py:92-98: forward + iou_score (host read of two counts)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / steps * 1e3
    del model
    torch.cuda.empty_cache()
    return {"images_per_s": n / ms * 1e3, "ms_per_batch": ms, "iou": iou}


res = {}
res["C1 NestedUNet woDS N=16 96x96 train"] = train_rate("NestedUNet", 1, False, 16, 96)
res["C3 UNet N=64 256x256 train"] = train_rate("UNet", 1, False, 64, 256)
res["C3 NestedUNet woDS N=64 256x256 train"] = train_rate("NestedUNet", 1, False, 64, 256)
res["C4 NestedUNet 4-class DS N=8 512x512 train"] = train_rate("NestedUNet", 4, True, 8, 512, steps=3, warm=3)
for n in (1, 4, 16, 64, 256, 1024):
    res["C5 NestedUNet eval+IoU N=%d 96x96" % n] = eval_rate("NestedUNet", n, 96)
for n in (1, 4, 16, 64):
    res["C5 NestedUNet eval+IoU N=%d 256x256" % n] = eval_rate("NestedUNet", n, 256)
for k, v in res.items():
    print("%-50s %s" % (k, json.dumps(v)))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
