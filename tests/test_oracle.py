"""The CPU oracle (oracle/nunet_oracle.py) against goldens minted from the UNMODIFIED reference
(tests/golden/make_golden.py, run in the build container where /root/reference exists)."""
import os

import pytest
import torch

from oracle import nunet_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def steps():
    return torch.load(os.path.join(GOLD, "steps.pt"))


def _model_sd(g):
    import archs
    torch.manual_seed(g["weight_seed"])
    m = getattr(archs, g["arch"])(g["num_classes"], 3, g["deep_supervision"])  # parameter container; same init order as the reference
    return {k: v.clone() for k, v in m.state_dict().items()}


@pytest.mark.parametrize("name", ["nested_ds_n2_32", "nested_wods_n2_32_dsb", "unet_n4_64", "nested_ds_c4_n2_32", "nested_ds_n3_48x80"])
def test_oracle_train_step_matches_reference(steps, name):
    g = steps[name]
    sd = _model_sd(g)
    x, t = O.synthetic_batch(g["n"], g["h"], g["w"], g["num_classes"], 3, g["data_seed"], g["dist"])
    new_sd, loss, iou, grads, outs = O.train_step(sd, x, t, g["arch"], g["deep_supervision"])
    refs = g["logits"] if isinstance(g["logits"], list) else [g["logits"]]
    outs = outs if isinstance(outs, list) else [outs]
    for o, r in zip(outs, refs):
        assert torch.allclose(o, r, rtol=1e-4, atol=2e-5 * float(r.abs().max())), float((o - r).abs().max())
    assert abs(loss - g["loss"]) < 2e-6
    assert (O.iou_counts(outs[-1], t)) == (g["I"], g["U"])
    assert iou == g["iou"]
    for k, n_ref in g["grad_norms"].items():
        if k.endswith(("conv1.bias", "conv2.bias")):
            assert n_ref < 1e-4 and float(grads[k].norm()) < 1e-4   # analytically zero under train-mode BN: fp32 noise only
            continue
        assert abs(float(grads[k].norm()) - n_ref) <= 2e-3 * n_ref + 1e-7, k
    for k, g_ref in g["grads"].items():
        if k.endswith(("conv1.bias", "conv2.bias")):
            continue
        assert torch.allclose(grads[k], g_ref, rtol=1e-3, atol=2e-3 * float(g_ref.abs().max()) + 1e-9), k
    for k, v in g["bn_after"].items():
        assert torch.allclose(new_sd[k], v, rtol=1e-4, atol=1e-6), k
    for k, n_ref in g["param_after_norms"].items():
        assert abs(float(new_sd[k].float().norm()) - n_ref) <= 1e-5 * max(n_ref, 1.0), k
    # eval path with the post-step state
    le, ie, oe = O.eval_step(new_sd, x, t, g["arch"], g["deep_supervision"])
    oe_last = oe[-1] if isinstance(oe, list) else oe
    assert torch.allclose(oe_last, g["eval_logits_last"], rtol=1e-3, atol=1e-4 * float(g["eval_logits_last"].abs().max()))
    assert abs(le - g["eval_loss"]) < 1e-5 and abs(ie - g["eval_iou"]) < 1e-3


def test_oracle_primitives_match_reference():
    p = torch.load(os.path.join(GOLD, "primitives.pt"))
    assert torch.allclose(O.upsample_bilinear_2x(p["x"]), p["up"], rtol=1e-6, atol=1e-6)
    assert torch.equal(O.max_pool_2x2(p["x"]), p["pool"])


@pytest.mark.parametrize("name", ["hard_1c", "soft_1c", "hard_4c"])
def test_oracle_loss_and_iou_match_reference(name):
    g = torch.load(os.path.join(GOLD, "loss_iou.pt"))[name]
    x = g["logits"].clone().requires_grad_(True)
    loss = O.bce_dice_loss(x, g["target"])
    loss.backward()
    assert abs(float(loss) - g["loss"]) < 1e-6
    assert torch.allclose(x.grad, g["dlogits"], rtol=1e-4, atol=1e-7)
    assert O.iou_counts(g["logits"], g["target"]) == (g["I"], g["U"])
    assert O.iou_score(g["logits"], g["target"]) == g["iou"]


def test_graphs_have_reference_topology():
    ng = dict(O.nested_graph())
    assert list(ng)[:4] == ["0_0", "1_0", "0_1", "2_0"] and len(ng) == 15
    assert ng["0_4"] == [("0_0", "same"), ("0_1", "same"), ("0_2", "same"), ("0_3", "same"), ("1_3", "up")]
    assert ng["3_0"] == [("2_0", "pool")]
    ug = dict(O.unet_graph())
    assert len(ug) == 9 and ug["2_2"] == [("2_0", "same"), ("3_1", "up")]
