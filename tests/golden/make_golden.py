"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU fp32.

Run in the build container only (the reference checkout does not exist on the GPU box):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
Weights are NOT stored (36 MB): they are re-created from torch.manual_seed(weight_seed) + default nn init, which the
reference modules and pytorch_nested-unet_b200/archs.py consume in the same order.  Inputs come from
oracle.nunet_oracle.synthetic_batch(seed).  Stored: logits, loss, IoU, integer I/U counts, selected gradients,
per-parameter gradient norms, post-step BatchNorm running statistics.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import nunet_oracle as O  # noqa: E402  (synthetic data generator only)


def load_ref():
    spec = importlib.util.spec_from_file_location("ref_archs1", os.path.join(REF, "finished", "archs1.py"))
    archs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(archs)
    import losses as ref_losses
    import metrics as ref_metrics
    return archs, ref_losses, ref_metrics


FULL_GRADS = ("final.weight", "final.bias", "final1.weight", "final1.bias", "final4.weight", "final4.bias",
              "conv0_0.conv1.weight", "conv0_0.bn1.weight", "conv0_0.bn1.bias", "conv0_4.conv2.weight", "conv0_4.bn2.weight",
              "conv1_0.conv1.weight", "conv4_0.bn2.bias", "conv0_1.conv1.weight", "conv3_1.bn1.weight")


def train_case(archs, losses, metrics, arch, num_classes, ds, n, h, w, dist, weight_seed=0, data_seed=1, warm=0):
    torch.manual_seed(weight_seed)
    model = getattr(archs, arch)(num_classes, 3, ds)
    crit = losses.BCEDiceLoss()
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
    x, t = O.synthetic_batch(n, h, w, num_classes, 3, data_seed, dist)
    model.train()
    for i in range(warm):               # `warm` full SGD steps of the reference loop body (trains.py:113-136) on other batches first
        xw, tw = O.synthetic_batch(n, h, w, num_classes, 3, data_seed + 1000 + i, dist)
        ow = model(xw)
        lw = sum(crit(o, tw) for o in ow) / len(ow) if isinstance(ow, list) else crit(ow, tw)
        opt.zero_grad()
        lw.backward()
        opt.step()
    out = model(x)
    if isinstance(out, list):
        loss = 0
        for o in out:
            loss = loss + crit(o, t)
        loss = loss / len(out)
        last = out[-1]
    else:
        loss = crit(out, t)
        last = out
    iou = metrics.iou_score(last, t)
    opt.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    opt.step()
    sd_after = model.state_dict()
    o_ = torch.sigmoid(last).detach().numpy() > 0.5
    t_ = t.numpy() > 0.5
    rec = {
        "arch": arch, "num_classes": num_classes, "deep_supervision": ds, "n": n, "h": h, "w": w, "dist": dist,
        "weight_seed": weight_seed, "data_seed": data_seed, "warm": warm,
        "logits": [o.detach().clone() for o in out] if isinstance(out, list) else out.detach().clone(),
        "loss": float(loss.detach()), "iou": float(iou),
        "I": int((o_ & t_).sum()), "U": int((o_ | t_).sum()),
        "grad_norms": {k: float(g.norm()) for k, g in grads.items()},
        "grads": {k: grads[k] for k in FULL_GRADS if k in grads},
        "bn_after": {k: v.clone() for k, v in sd_after.items() if "running" in k},
        "param_after_norms": {k: float(v.float().norm()) for k, v in sd_after.items() if v.is_floating_point()},
    }
    # eval-mode forward with the post-step weights and running stats (val.py path)
    model.eval()
    with torch.no_grad():
        oe = model(x)
        oe_last = oe[-1] if isinstance(oe, list) else oe
        rec["eval_logits_last"] = oe_last.clone()
        rec["eval_iou"] = float(metrics.iou_score(oe_last, t))
        if isinstance(oe, list):
            le = sum(crit(o, t) for o in oe) / len(oe)
        else:
            le = crit(oe, t)
        rec["eval_loss"] = float(le)
    return rec


def loss_case(losses, metrics, seed, shape, soft):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g) * 3
    flat = x.view(-1)
    flat[::7] = 0.0                      # exact zeros: sigmoid(0) = 0.5 is NOT > 0.5
    flat[1::11] = 5e-8                   # below the fp32 sigmoid threshold (~8.94e-8): still False
    flat[2::13] = 1e-6                   # above it: True
    t = torch.rand(shape, generator=g)
    if not soft:
        t = (t > 0.5).float()
    xr = x.clone().requires_grad_(True)
    loss = losses.BCEDiceLoss()(xr, t)
    loss.backward()
    o_ = torch.sigmoid(x).numpy() > 0.5
    t_ = t.numpy() > 0.5
    return {"logits": x, "target": t, "loss": float(loss.detach()), "dlogits": xr.grad.clone(),
            "iou": float(metrics.iou_score(x, t)), "I": int((o_ & t_).sum()), "U": int((o_ | t_).sum())}


def main():
    archs, losses, metrics = load_ref()
    out = {}
    out["nested_ds_n2_32"] = train_case(archs, losses, metrics, "NestedUNet", 1, True, 2, 32, 32, "normal")
    out["nested_wods_n2_32_dsb"] = train_case(archs, losses, metrics, "NestedUNet", 1, False, 2, 32, 32, "dsb")
    out["unet_n4_64"] = train_case(archs, losses, metrics, "UNet", 1, False, 4, 64, 64, "normal")
    out["nested_ds_c4_n2_32"] = train_case(archs, losses, metrics, "NestedUNet", 4, True, 2, 32, 32, "normal")
    out["nested_wods_n4_96"] = train_case(archs, losses, metrics, "NestedUNet", 1, False, 4, 96, 96, "normal")
    out["nested_ds_n3_48x80"] = train_case(archs, losses, metrics, "NestedUNet", 1, True, 3, 48, 80, "normal", weight_seed=3, data_seed=7)
    torch.save(out, os.path.join(ROOT, "tests", "golden", "steps.pt"))
    lo = {
        "hard_1c": loss_case(losses, metrics, 5, (3, 1, 16, 16), False),
        "soft_1c": loss_case(losses, metrics, 6, (2, 1, 32, 16), True),
        "hard_4c": loss_case(losses, metrics, 7, (2, 4, 16, 32), False),
    }
    torch.save(lo, os.path.join(ROOT, "tests", "golden", "loss_iou.pt"))
    # primitive ops of the skip pathway
    g = torch.Generator().manual_seed(11)
    xs = torch.randn(2, 5, 6, 10, generator=g)
    prim = {"x": xs,
            "up": torch.nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)(xs),
            "pool": torch.nn.MaxPool2d(2, 2)(xs)}
    torch.save(prim, os.path.join(ROOT, "tests", "golden", "primitives.pt"))
    for f in ("steps.pt", "loss_iou.pt", "primitives.pt"):
        print(f, os.path.getsize(os.path.join(ROOT, "tests", "golden", f)))
    for k, v in out.items():
        print(k, "loss", v["loss"], "iou", v["iou"], "I", v["I"], "U", v["U"])


if __name__ == "__main__":
    main()
