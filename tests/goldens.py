"""Shared access to the reference-minted goldens (tests/golden/*.pt) and reconstruction of the model state a case starts from."""
import os

import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_CACHE = {}


def load(fname):
    if fname not in _CACHE:
        _CACHE[fname] = torch.load(os.path.join(GOLD, fname))
    return _CACHE[fname]


def case(name):
    for f in ("steps.pt", "steps_r02.pt"):
        d = load(f)
        if name in d:
            return d[name]
    raise KeyError(name)


def autocast_context(name):
    """Per-head error of the unmodified reference under torch.autocast(bfloat16) against its own fp32 logits (or None)."""
    return load("steps_r02.pt")["_autocast_bf16"].get(name)


def autocast_grads(name):
    """{parameter: (relative deviation of the gradient norm, cosine)} of the unmodified reference's bf16-autocast training step against
    its own fp32 gradients (tests/golden/make_autocast_grads.py), or None."""
    return load("autocast_grads.pt").get(name)


def synthetic(g, offset=0):
    from oracle.nunet_oracle import synthetic_batch
    return synthetic_batch(g["n"], g["h"], g["w"], g["num_classes"], 3, g["data_seed"] + offset, g["dist"])


def start_state(g):
    """(state_dict, SGD momentum buffers) the recorded step of golden `g` starts from: default init under the weight seed, then — for a
    warm case — g['warm'] SGD steps of the oracle's restatement of trains.py:113-136 on the same batches make_golden.train_case used.
    (Weights are not stored in the fixtures: 36 MB per case.)"""
    import archs
    from oracle import nunet_oracle as O
    torch.manual_seed(g["weight_seed"])
    m = getattr(archs, g["arch"])(g["num_classes"], 3, g["deep_supervision"])  # parameter container; same init order as the reference
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    bufs = {}
    for i in range(g.get("warm", 0)):
        xw, tw = synthetic(g, 1000 + i)
        sd, _, _, _, _ = O.train_step(sd, xw, tw, g["arch"], g["deep_supervision"], bufs=bufs)
    return sd, bufs


def grad_tolerances(name, floor=1e-3, factor=4.0):
    """(norm tolerance, rel-L2 tolerance) for comparing gradients with the reference's: `floor`, or `factor` x how far the reference's own
    gradients move under a 1e-7 relative input perturbation (discontinuous ReLU gates / max-pool winners; recorded per case by
    tests/golden/make_golden_r02.py), whichever is larger."""
    s = load("steps_r02.pt")["_grad_sensitivity"].get(name)
    if s is None:
        return floor, floor
    return max(floor, factor * s["worst_norm_rel"]), max(floor, factor * max(s["rell2"].values()))
