"""Round-2 goldens at the BASELINE shapes (tests/golden/steps_r02.pt), minted from the UNMODIFIED reference exactly like
make_golden.py (same train_case; run in the build container only):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_r02.py
  c1_nested_wods_n16_96_dsb   BASELINE config 1 exactly: NestedUNet woDS, 3ch -> 1 class, 96x96, batch 16, loader distribution
  c3_nested_wods_n2_256       ISIC shape (train_ISIC.py has no deep supervision), 256x256: rows wider than the 254-px TMA box
  nested_ds_c4_n2_32x272      config-4-like: 4 classes, deep supervision, W = 272 > 254 (mixed materialised / fused-loader engine branch)
  unet_n2_32x272              the same width limit through archs.UNet
  warm_nested_ds_n4_32        20 reference SGD steps (momentum 0.9, wd 1e-4) on other batches, then the recorded step: weights,
                              momentum buffers and BatchNorm running statistics have left their initial values
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


def autocast_error(archs, g):
    torch.manual_seed(g["weight_seed"])
    model = getattr(archs, g["arch"])(g["num_classes"], 3, g["deep_supervision"])
    x, _ = G.O.synthetic_batch(g["n"], g["h"], g["w"], g["num_classes"], 3, g["data_seed"], g["dist"])
    if g.get("warm"):
        return {"relmax": [], "rell2": []}          # the warmed weights are not stored; see tests/goldens.py
    model.train()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        o = model(x)
    o = o if isinstance(o, list) else [o]
    refs = g["logits"] if isinstance(g["logits"], list) else [g["logits"]]
    return {"relmax": [float((a.float() - r).abs().max() / r.abs().max()) for a, r in zip(o, refs)],
            "rell2": [float((a.float() - r).norm() / r.norm()) for a, r in zip(o, refs)]}


def grad_sensitivity(archs, losses, g):
    def grads_of(scale):
        torch.manual_seed(g["weight_seed"])
        model = getattr(archs, g["arch"])(g["num_classes"], 3, g["deep_supervision"])
        crit = losses.BCEDiceLoss()
        x, t = G.O.synthetic_batch(g["n"], g["h"], g["w"], g["num_classes"], 3, g["data_seed"], g["dist"])
        model.train()
        o = model(x * scale)
        loss = sum(crit(a, t) for a in o) / len(o) if isinstance(o, list) else crit(o, t)
        loss.backward()
        return {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    x0, _ = G.O.synthetic_batch(g["n"], g["h"], g["w"], g["num_classes"], 3, g["data_seed"], g["dist"])
    pert = 1 + 1e-7 * torch.randn(x0.shape, generator=torch.Generator().manual_seed(123))
    a, b = grads_of(1.0), grads_of(pert)
    skip = ("conv1.bias", "conv2.bias")
    return {"rell2": {k: float((b[k] - a[k]).norm() / a[k].norm()) for k in G.FULL_GRADS if k in a and not k.endswith(skip)},
            "worst_norm_rel": max(abs(float(b[k].norm()) - float(a[k].norm())) / float(a[k].norm()) for k in a if not k.endswith(skip))}


def checkpoint_interchange(archs):
    """model.pth interchange (trains.py:344-346 / val.py:58-59), recorded from the REFERENCE side:
      * manifest: key -> (shape, dtype) of the state_dict the unmodified reference modules save, in order;
      * the reference modules load (strict) a state_dict saved by this repository's archs modules;
      * logits of the reference in eval mode on a checkpoint written HERE (values are re-created from the seed in the test)."""
    import io
    sys.path.insert(0, os.path.join(G.ROOT, "pytorch_nested-unet_b200"))
    import archs as ours
    rec = {}
    for name, args in (("NestedUNet_ds", ("NestedUNet", 1, 3, True)), ("NestedUNet", ("NestedUNet", 1, 3, False)), ("UNet", ("UNet", 1, 3, False)),
                       ("NestedUNet_c4_ds", ("NestedUNet", 4, 3, True))):
        arch, k, c, ds = args
        torch.manual_seed(7)
        ref = getattr(archs, arch)(k, c, ds)
        buf = io.BytesIO()
        torch.save(ref.state_dict(), buf)          # what trains.py:345 writes
        buf.seek(0)
        ref_sd = torch.load(buf)
        torch.manual_seed(7)
        mine = getattr(ours, arch)(k, c, ds)
        mine.load_state_dict(ref_sd, strict=True)   # the reference's checkpoint loads here ...
        buf2 = io.BytesIO()
        torch.save(mine.state_dict(), buf2)
        buf2.seek(0)
        ref2 = getattr(archs, arch)(k, c, ds)
        ref2.load_state_dict(torch.load(buf2), strict=True)   # ... and a checkpoint written here loads in the reference
        same = all(torch.equal(a, b) for a, b in zip(ref.state_dict().values(), ref2.state_dict().values()))
        rec[name] = {"manifest": [(key, tuple(v.shape), str(v.dtype)) for key, v in ref_sd.items()], "round_trip_equal": bool(same)}
    return rec


def main():
    archs, losses, metrics = G.load_ref()
    out = {}
    out["c1_nested_wods_n16_96_dsb"] = G.train_case(archs, losses, metrics, "NestedUNet", 1, False, 16, 96, 96, "dsb")
    out["c3_nested_wods_n2_256"] = G.train_case(archs, losses, metrics, "NestedUNet", 1, False, 2, 256, 256, "normal", weight_seed=2, data_seed=5)
    out["nested_ds_c4_n2_32x272"] = G.train_case(archs, losses, metrics, "NestedUNet", 4, True, 2, 32, 272, "normal", weight_seed=4, data_seed=9)
    out["unet_n2_32x272"] = G.train_case(archs, losses, metrics, "UNet", 1, False, 2, 32, 272, "normal", weight_seed=5, data_seed=10)
    out["warm_nested_ds_n4_32"] = G.train_case(archs, losses, metrics, "NestedUNet", 1, True, 4, 32, 32, "normal", weight_seed=6, data_seed=11, warm=20)
    # context for the bf16 tolerances: the UNMODIFIED reference under torch.autocast(bfloat16) against its own fp32 logits (train mode)
    small = G.torch.load(os.path.join(G.ROOT, "tests", "golden", "steps.pt"))
    ac = {}
    for name, g in list(small.items()) + list(out.items()):
        if g["n"] * g["h"] * g["w"] > 4 * 96 * 96:
            continue
        ac[name] = autocast_error(archs, g)
        print(name, "autocast bf16 relmax per head", ["%.2e" % v for v in ac[name]["relmax"]])
    out["_autocast_bf16"] = ac
    # how far the REFERENCE's OWN gradients move when its input is perturbed by 1e-7 (relative): ReLU gates and max-pool winners are
    # discontinuous, and with N*H*W in the 10^5 range a handful of near-ties flip, so gradients of two fp32 evaluations that agree to
    # 1e-6 in the logits differ by up to ~5e-3 (rel-L2).  This is the floor any non-bitwise-identical implementation can be held to.
    sens = {}
    for name, g in list(small.items()) + list(out.items()):
        if name.startswith("_") or g.get("warm"):
            continue
        sens[name] = grad_sensitivity(archs, losses, g)
        print(name, "gradient sensitivity to a 1e-7 input perturbation: worst rel-L2 %.2e, worst norm %.2e" % (
            max(sens[name]["rell2"].values()), sens[name]["worst_norm_rel"]))
    out["_grad_sensitivity"] = sens
    out["_checkpoints"] = checkpoint_interchange(archs)
    lov = {}
    for name, seed, shape in (("n3_16", 31, (3, 1, 16, 16)), ("n2_96", 32, (2, 1, 96, 96)), ("c4_n2", 33, (2, 4, 16, 32)), ("n1_160", 34, (1, 1, 160, 160))):
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(shape, generator=g) * 2
        t = (torch.rand(shape, generator=g) > 0.6).float()
        if name == "n3_16":
            t[1] = 0.0                 # an image without foreground
        xr = x.clone().requires_grad_(True)
        loss = losses.LovaszHingeLoss()(xr, t)
        loss.backward()
        lov[name] = {"logits": x, "target": t.to(torch.uint8), "loss": float(loss.detach()), "dlogits": xr.grad.clone()}
        print("lovasz", name, lov[name]["loss"])
    out["_lovasz"] = lov
    path = os.path.join(G.ROOT, "tests", "golden", "steps_r02.pt")
    torch.save(out, path)
    print("steps_r02.pt", os.path.getsize(path))
    for k, v in out.items():
        if not k.startswith("_"):
            print(k, "loss", v["loss"], "iou", v["iou"], "I", v["I"], "U", v["U"])


if __name__ == "__main__":
    main()
