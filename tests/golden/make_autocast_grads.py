"""Context for the bf16 GRADIENT tolerances (tests/golden/autocast_grads.pt), minted from the UNMODIFIED reference (build container):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_autocast_grads.py
For every small golden case: the reference's own training step under torch.autocast(bfloat16) (forward AND backward) against its own
fp32 gradients — per parameter the relative deviation of the gradient norm and the cosine.  Some BatchNorm gradients of the loader-
distribution cases are badly conditioned (sums of g*yhat that cancel): the bf16 product is held to a multiple of what the reference
itself loses there, not to a hand-picked constant."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


def grads_of(archs, losses, g, autocast):
    torch.manual_seed(g["weight_seed"])
    model = getattr(archs, g["arch"])(g["num_classes"], 3, g["deep_supervision"])
    crit = losses.BCEDiceLoss()
    x, t = G.O.synthetic_batch(g["n"], g["h"], g["w"], g["num_classes"], 3, g["data_seed"], g["dist"])
    model.train()
    if autocast:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            o = model(x)
    else:
        o = model(x)
    o = [a.float() for a in o] if isinstance(o, list) else o.float()
    loss = sum(crit(a, t) for a in o) / len(o) if isinstance(o, list) else crit(o, t)
    loss.backward()
    return {k: p.grad.detach().float().clone() for k, p in model.named_parameters()}


def main():
    archs, losses, metrics = G.load_ref()
    cases = dict(torch.load(os.path.join(G.ROOT, "tests", "golden", "steps.pt")))
    cases.update({k: v for k, v in torch.load(os.path.join(G.ROOT, "tests", "golden", "steps_r02.pt")).items() if not k.startswith("_")})
    out = {}
    for name, g in cases.items():
        if g.get("warm") or g["n"] * g["h"] * g["w"] > 16 * 96 * 96:
            continue
        a, b = grads_of(archs, losses, g, False), grads_of(archs, losses, g, True)
        rec = {}
        for k in a:
            if k.endswith(("conv1.bias", "conv2.bias")):
                continue
            na, nb = float(a[k].norm()), float(b[k].norm())
            rec[k] = (abs(nb - na) / max(na, 1e-12), float((a[k] * b[k]).sum() / (na * nb + 1e-30)))
        out[name] = rec
        worst = max(rec.items(), key=lambda kv: kv[1][0])
        print(name, "worst norm deviation under autocast", worst[0], "%.3f cos %.3f" % worst[1], "| lowest cos %.3f" % min(v[1] for v in rec.values()))
    torch.save(out, os.path.join(HERE, "autocast_grads.pt"))


if __name__ == "__main__":
    main()
