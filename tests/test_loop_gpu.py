"""The reference's scripts end to end on the GPU: train.main() (trains.py:191-356) on synthetic data writes
models/<name>/config.yml, log.csv and the best-IoU model.pth; val.main() (val.py:31-109) reloads them, prints the mean IoU and writes
one uint8 plane per image and class.  Also: a VGGBlock on its own, the val output stage kernel, eval graphs across shapes."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("arch,ds,opt,u8", [("NestedUNet", "True", "SGD", "True"), ("UNet", "False", "Adam", "False")])
def test_train_main_then_val_main(tmp_path, monkeypatch, arch, ds, opt, u8):
    import train
    import val
    import yaml
    monkeypatch.chdir(tmp_path)
    name = "loop_%s" % arch
    log = train.main(["--name", name, "--dataset", "synthetic", "--arch", arch, "--deep_supervision", ds, "--epochs", "3", "-b", "4",
                      "--input_w", "32", "--input_h", "32", "--synthetic_batches", "4", "--optimizer", opt, "--lr", "0.05" if opt == "SGD" else "0.002",
                      "--u8_input", u8])
    d = tmp_path / "models" / name
    assert (d / "config.yml").exists() and (d / "log.csv").exists() and (d / "model.pth").exists()
    cfg = yaml.safe_load(open(d / "config.yml"))
    assert cfg["arch"] == arch and cfg["batch_size"] == 4 and cfg["optimizer"] == opt
    rows = open(d / "log.csv").read().strip().split("\n")
    assert rows[0] == ",epoch,lr,loss,iou,val_loss,val_iou" and len(rows) == 4
    assert all(0.0 < v < 5.0 for v in log["loss"]) and all(0.0 <= v <= 1.0 for v in log["val_iou"])
    assert log["loss"][-1] < log["loss"][0]                     # it learns the (image-correlated) synthetic masks
    sd = torch.load(d / "model.pth")
    assert "conv0_0.conv1.weight" in sd and sd["conv0_0.bn1.num_batches_tracked"] > 0
    iou = val.main(["--name", name])
    assert 0.0 <= iou <= 1.0
    outs = sorted(os.listdir(tmp_path / "outputs" / name / "0"))
    assert len(outs) == 4 and outs[0].startswith("syn0_0")       # one plane per validation image (1 batch of 4)
    assert abs(iou - max(log["val_iou"])) < 0.2                  # the reloaded best checkpoint scores like the epoch that saved it


def test_sigmoid_u8_matches_numpy_semantics():
    import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 2, 17, 9, generator=g) * 4
    x.view(-1)[::5] = 0.0
    got = ops.sigmoid_u8(x.to(DEV)).cpu().numpy()
    ref = (torch.sigmoid(x).numpy() * 255).astype("uint8")        # val.py:100,105
    diff = (got.astype(int) - ref.astype(int))
    assert abs(diff).max() <= 1 and (diff != 0).mean() < 0.01   # expf vs the host's sigmoid: a value exactly on a step may land either side


@pytest.mark.parametrize("cin,cmid,cout,train", [(3, 32, 32, True), (96, 32, 32, True), (64, 64, 64, False), (40, 64, 32, True)])
def test_vgg_block_on_its_own(cin, cmid, cout, train):
    """archs.VGGBlock.forward / backward (finished/archs1.py:23-32) against torch's nn.Module composition."""
    import archs
    from torch import nn
    torch.manual_seed(3)
    blk = archs.VGGBlock(cin, cmid, cout).to(DEV)
    ref = nn.Sequential(nn.Conv2d(cin, cmid, 3, padding=1), nn.BatchNorm2d(cmid), nn.ReLU(), nn.Conv2d(cmid, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU()).to(DEV)
    ref[0].load_state_dict(blk.conv1.state_dict()); ref[1].load_state_dict(blk.bn1.state_dict())
    ref[3].load_state_dict(blk.conv2.state_dict()); ref[4].load_state_dict(blk.bn2.state_dict())
    blk.train(train); ref.train(train)
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn(3, cin, 16, 24, generator=g, device=DEV)
    xr = x.clone().requires_grad_(True)
    xo = x.clone().requires_grad_(True)
    o, r = blk(xo), ref(xr)
    assert o.shape == r.shape and float((o - r).abs().max()) <= 3e-2 * float(r.abs().max())
    if train:
        w = torch.randn_like(r)
        (o * w).sum().backward()
        (r * w).sum().backward()
        for a, b in ((blk.conv1.weight.grad, ref[0].weight.grad), (blk.bn2.weight.grad, ref[4].weight.grad), (blk.conv2.weight.grad, ref[3].weight.grad),
                     (xo.grad, xr.grad)):
            assert float((a - b).norm() / b.norm()) <= 8e-2
        assert torch.allclose(blk.bn1.running_mean, ref[1].running_mean, rtol=2e-2, atol=2e-3)


def test_eval_graphs_of_different_shapes_do_not_share_weight_images():
    """eval at shape A, then B (different tiling of the same weights), then A again must replay A's graph over ITS images."""
    import archs
    torch.manual_seed(0)
    m = archs.NestedUNet(1, 3, False).to(DEV).eval()
    g = torch.Generator(device=DEV).manual_seed(2)
    xa = torch.randn(2, 3, 32, 32, generator=g, device=DEV)
    xb = torch.randn(1, 3, 64, 288, generator=g, device=DEV)      # rows wider than 254 px: other loaders, other tilings
    with torch.no_grad():
        a1 = m(xa).clone()
        m(xb)
        a2 = m(xa).clone()
        m(xb)
        a3 = m(xa).clone()
    assert torch.equal(a1, a2) and torch.equal(a1, a3)


def test_step_scalars_and_prefetcher_give_the_synchronous_loop_values():
    """utils.StepScalars / utils.DevicePrefetcher (train.train's loop): the scalars collected after the step equal loss.item() and
    metrics.iou_score of the same step, batches arrive unchanged and in order, keep=k leaves k submissions in flight."""
    import archs
    import losses
    import metrics
    import utils
    torch.manual_seed(5)
    model = archs.NestedUNet(1, 3, True).to(DEV).train()
    crit = losses.BCEDiceLoss()
    g = torch.Generator().manual_seed(2)
    host = [((torch.randn(2, 3, 32, 32, generator=g) * 0.01).pin_memory(), (torch.rand(2, 1, 32, 32, generator=g) > 0.5).float().pin_memory(), {"i": k})
            for k in range(4)]
    scal = utils.StepScalars()
    seen = []
    for k, (x, t, meta) in enumerate(utils.DevicePrefetcher(iter(host), DEV)):
        assert meta["i"] == k and x.is_cuda and torch.equal(x.cpu(), host[k][0]) and torch.equal(t.cpu(), host[k][1])
        outs = model(x)
        loss = sum(crit(o, t) for o in outs) / len(outs)
        counts = metrics.iou_counts_device(outs[-1], t)
        scal.submit(loss, counts)
        loss.backward()
        got = scal.collect(keep=1 if k < 2 else 0)
        seen += got
        ref_iou = metrics.iou_score(outs[-1], t)
        if k == 3:
            lv, i_, u_ = got[-1]
            assert abs(lv - loss.item()) == 0.0 and abs((i_ + 1e-5) / (u_ + 1e-5) - ref_iou) == 0.0
    assert len(seen) == 4 and not scal.slots
    # counts without the loss cache: one counting launch, same integers
    o = torch.randn(2, 1, 16, 16, device=DEV)
    t = (torch.rand(2, 1, 16, 16, device=DEV) > 0.5).float()
    assert tuple(metrics.iou_counts_device(o, t).tolist()) == metrics.iou_counts(o, t)
