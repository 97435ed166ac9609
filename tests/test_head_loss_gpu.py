"""Heads, BCE+Dice loss, IoU counts: CUDA kernels vs the committed reference goldens and vs torch on the device."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["hard_1c", "soft_1c", "hard_4c"])
def test_loss_iou_against_reference_golden(name):
    import ops
    g = torch.load(os.path.join(GOLD, "loss_iou.pt"))[name]
    x, t = g["logits"].to(DEV), g["target"].to(DEV)
    loss, sums, counts = ops.loss_iou_fwd(x, t)
    gs = torch.tensor(1.0, device=DEV)
    dl = ops.loss_bwd(x, t, sums, gs)
    c2 = ops.iou_counts(x, t)
    torch.cuda.synchronize()
    assert abs(float(loss) - g["loss"]) <= 2e-6 * max(1.0, abs(g["loss"]))
    assert counts.tolist() == [g["I"], g["U"]]          # bit-exact integer counts (incl. logit == 0 -> False)
    assert c2.tolist() == [g["I"], g["U"]]
    assert (c2[0].item() + 1e-5) / (c2[1].item() + 1e-5) == g["iou"]
    ref = g["dlogits"].to(DEV)
    assert torch.allclose(dl, ref, rtol=1e-4, atol=1e-7 + 1e-5 * float(ref.abs().max()))


def test_loss_bwd_scale_pointer_and_mul():
    import ops
    gen = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(4, 1, 96, 96, generator=gen, device=DEV)
    t = (torch.rand(4, 1, 96, 96, generator=gen, device=DEV) > 0.5).float()
    loss, sums, counts = ops.loss_iou_fwd(x, t)
    d1 = ops.loss_bwd(x, t, sums, torch.tensor(0.5, device=DEV), 0.25)
    xr = x.clone().requires_grad_(True)
    bce = F.binary_cross_entropy_with_logits(xr, t)
    p = torch.sigmoid(xr).view(4, -1)
    tt = t.view(4, -1)
    dice = (2 * (p * tt).sum(1) + 1e-5) / (p.sum(1) + tt.sum(1) + 1e-5)
    ref_loss = 0.5 * bce + 1 - dice.sum() / 4
    ref_loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref_loss)) < 2e-6
    assert torch.allclose(d1, xr.grad * 0.125, rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize("K", [1, 4])
def test_head_fwd_and_param_grads(K):
    import ops
    N, H, W = 3, 16, 24
    gen = torch.Generator(device=DEV).manual_seed(2)
    y = torch.randn(N, H, W, 32, generator=gen, device=DEV).to(torch.bfloat16)
    scale = torch.randn(32, generator=gen, device=DEV)
    shift = torch.randn(32, generator=gen, device=DEV) * 0.2
    wh = torch.randn(K, 32, generator=gen, device=DEV) * 0.3
    bh = torch.randn(K, generator=gen, device=DEV)
    logits = ops.head_fwd(y, scale, shift, wh, bh, N, H, W)
    a = torch.relu(y.float().permute(0, 3, 1, 2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    ref = F.conv2d(a, wh.view(K, 32, 1, 1), bh)
    torch.cuda.synchronize()
    assert torch.allclose(logits, ref, rtol=1e-5, atol=1e-5)
    dl = torch.randn(N, K, H, W, generator=gen, device=DEV)
    dwh, dbh = ops.head_bwd_params(y, scale, shift, dl, N, H, W, K)
    torch.cuda.synchronize()
    rw = torch.einsum("nkhw,nchw->kc", dl, a)
    rb = dl.sum((0, 2, 3))
    assert torch.allclose(dwh, rw, rtol=1e-4, atol=1e-3)
    assert torch.allclose(dbh, rb, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("name", ["n3_16", "n2_96", "c4_n2", "n1_160"])
def test_lovasz_hinge_loss_matches_reference_golden(name):
    """losses.LovaszHingeLoss (per-image sort + scan kernel, csrc/lovasz.cu) against loss and gradient minted from the reference's
    losses.LovaszHingeLoss (tests/golden/make_golden_r02.py): shared-memory sort (P <= 16384), an image without foreground, four
    classes flattened jointly, and the global-scratch sort (P = 25600)."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import goldens
    import losses
    g = goldens.load("steps_r02.pt")["_lovasz"][name]
    x = g["logits"].to("cuda:0").requires_grad_(True)
    t = g["target"].float().to("cuda:0")
    loss = losses.LovaszHingeLoss()(x, t)
    (loss * 2.0).backward()
    assert abs(float(loss.detach()) - g["loss"]) <= 2e-6 * max(1.0, abs(g["loss"])), (float(loss.detach()), g["loss"])
    ref = g["dlogits"] * 2.0
    d = x.grad.cpu()
    # ties between equal errors may be ordered differently than torch.sort did (the loss is invariant, the per-element subgradient is not)
    bad = (d - ref).abs() > 1e-6 * ref.abs().max() + 1e-9
    assert int(bad.sum()) <= 2, int(bad.sum())
    with torch.no_grad():
        assert abs(float(losses.LovaszHingeLoss()(x.detach(), t)) - g["loss"]) <= 2e-6 * max(1.0, abs(g["loss"]))


@pytest.mark.parametrize("nh,K", [(4, 1), (2, 4), (3, 1)])
def test_multi_head_loss_is_bitwise_the_per_head_launches(nh, K):
    """nunet_loss_iou_fwd_multi / nunet_loss_bwd_multi (the deep-supervision loop of trains.py:120-123 in one pass) against
    nunet_loss_iou_fwd / nunet_loss_bwd head by head: identical bits, including a head whose loss received no gradient."""
    import ops
    gen = torch.Generator(device=DEV).manual_seed(10 + nh)
    xs = [torch.randn(5, K, 40, 56, generator=gen, device=DEV) * (1 + k) for k in range(nh)]
    xs[0].view(-1)[::7] = 0.0
    t = (torch.rand(5, K, 40, 56, generator=gen, device=DEV) > 0.5).float()
    loss, sums, counts = ops.loss_iou_fwd_multi(xs, t)
    gs = [torch.tensor(0.25 * (k + 1), device=DEV) for k in range(nh)]
    gs[1] = None
    dls = ops.loss_bwd_multi(xs, t, sums, gs, 0.5)
    assert dls[1] is None
    for k in range(nh):
        l1, s1, c1 = ops.loss_iou_fwd(xs[k], t)
        assert torch.equal(l1, loss[k]) and torch.equal(s1, sums[k]) and torch.equal(c1, counts[k])
        if gs[k] is not None:
            assert torch.equal(ops.loss_bwd(xs[k], t, s1, gs[k], 0.5), dls[k])


def test_grouped_criterion_matches_the_per_head_criterion():
    """losses.BCEDiceLoss on a marked group of heads (what NestedUNet(deep_supervision=True) returns): same losses, same cached IoU
    counts, same input gradients as calling it head by head on unmarked tensors; a partial loop and a second target still work."""
    import losses
    import metrics
    crit = losses.BCEDiceLoss()
    gen = torch.Generator(device=DEV).manual_seed(3)
    base = [torch.randn(4, 1, 32, 32, generator=gen, device=DEV) for _ in range(4)]
    t = (torch.rand(4, 1, 32, 32, generator=gen, device=DEV) > 0.5).float()
    t2 = (torch.rand(4, 1, 32, 32, generator=gen, device=DEV) > 0.5).float()
    a = [b.clone().requires_grad_(True) for b in base]
    r = [b.clone().requires_grad_(True) for b in base]
    heads = losses.mark_head_group([x * 1.0 for x in a])          # non-leaf tensors, as the engine's outputs are
    la = [crit(h, t) for h in heads]
    lr = [crit(x * 1.0, t) for x in r]
    for k in range(4):
        assert torch.equal(la[k].detach(), lr[k].detach())
        assert metrics.iou_counts(heads[k], t) == metrics.iou_counts(r[k].detach(), t)
    (sum(la) / 4).backward()
    (sum(lr) / 4).backward()
    for k in range(4):
        assert torch.equal(a[k].grad, r[k].grad)
    # only two of the heads enter the total; another target re-evaluates the group
    for x in a:
        x.grad = None
    heads = losses.mark_head_group([x * 1.0 for x in a])
    tot = crit(heads[0], t2) + crit(heads[3], t2)
    tot.backward()
    ref0 = base[0].clone().requires_grad_(True)
    crit(ref0 * 1.0, t2).backward()
    assert torch.equal(a[0].grad, ref0.grad) and float(a[1].grad.abs().max()) == 0.0 and float(a[3].grad.abs().max()) > 0.0
    with torch.no_grad():
        heads = losses.mark_head_group([b.clone() for b in base])
        assert torch.equal(crit(heads[2], t), lr[2].detach())
