"""Host-side logic that needs no GPU: engine graph bookkeeping, drop-in module surface, no-CPU-fallback behaviour,
gradient bucketing and the SyncBN statistics exchange over gloo with world_size 2."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_engine_specs_match_reference_graph():
    import archs
    import engine
    import _lib as L
    m = archs.NestedUNet(2, 3, True)
    specs = engine._state(m).specs
    assert len(specs) == 30
    by = {(s.node, s.which): s for s in specs}
    s = by[("0_4", 1)]
    assert s.cin_pad == 192 and s.cout == 32 and [k for _, k, _ in s.sources] == [L.SRC_SAME] * 4 + [L.SRC_UP]
    assert [c for _, _, c in s.sources] == [32, 32, 32, 32, 64]
    assert by[("0_0", 1)].cin_pad == 16 and by[("0_0", 1)].sources[0][1] == L.SRC_INPUT
    assert by[("3_0", 1)].sources == [(by[("2_0", 2)].idx, L.SRC_POOL, 128)]
    # x0_0 is consumed by conv1_0 (pool) and conv0_1..conv0_4 (same, channel offset 0)
    cons = by[("0_0", 2)].consumers
    assert sorted(c[2] for c in cons) == sorted([L.SRC_POOL] + [L.SRC_SAME] * 4) and all(c[1] == 0 for c in cons)
    # x1_2 enters conv0_3.conv1 upsampled at channel offset 3*32
    c13 = [c for c in by[("1_2", 2)].consumers if specs[c[0]].node == "0_3"]
    assert c13 == [(by[("0_3", 1)].idx, 96, L.SRC_UP, 3)]      # ... as the fourth source (position 3) of that conv
    assert [s.node for s in specs if s.head is not None] == ["0_1", "0_2", "0_3", "0_4"]
    u = archs.UNet(1, 3)
    assert len(engine._state(u).specs) == 18 and [s.node for s in engine._state(u).specs if s.head is not None] == ["0_4"]


def test_state_dict_layout_and_param_counts():
    import archs
    m = archs.NestedUNet(1, 3, True)
    keys = list(m.state_dict().keys())
    assert len(keys) == 218 and keys[0] == "conv0_0.conv1.weight"
    assert m.state_dict()["conv0_4.conv1.weight"].shape == (32, 192, 3, 3)
    assert sum(p.numel() for p in m.parameters()) == 9163428
    assert sum(p.numel() for p in archs.NestedUNet(1, 3, False).parameters()) == 9163329
    assert sum(p.numel() for p in archs.UNet(1, 3).parameters()) == 7852545
    assert archs.__all__ == ['UNet', 'NestedUNet']


def test_no_cpu_fallback():
    import archs
    import losses
    import _lib as L
    m = archs.UNet(1, 3)
    with pytest.raises(L.NunetError):
        m(torch.zeros(1, 3, 32, 32))
    with pytest.raises(L.NunetError):
        losses.BCEDiceLoss()(torch.zeros(1, 1, 8, 8), torch.zeros(1, 1, 8, 8))
    with pytest.raises(ValueError):
        if torch.cuda.is_available():
            m.cuda()(torch.zeros(1, 3, 40, 40, device="cuda"))
        else:
            raise ValueError("no gpu: shape check covered on the GPU box")


def test_utils_and_cli_defaults_match_reference():
    import train
    import utils
    a = utils.AverageMeter()
    a.update(2.0, 3)
    a.update(4.0, 1)
    assert a.avg == 2.5 and a.count == 4 and a.val == 4.0
    assert utils.str2bool("True") is True and utils.str2bool("false") is False
    cfg = vars(train.parse_args([]))
    ref = dict(epochs=100, batch_size=16, arch="NestedUNet", deep_supervision=False, input_channels=3, num_classes=1, input_w=96,
               input_h=96, loss="BCEDiceLoss", dataset="dsb2018_96", optimizer="SGD", lr=1e-3, momentum=0.9, weight_decay=1e-4,
               nesterov=False, scheduler="CosineAnnealingLR", min_lr=1e-5, factor=0.1, patience=2, milestones="1,2",
               early_stopping=-1, num_workers=4)
    for k, v in ref.items():
        assert cfg[k] == v, k


# ------------------------------------------------------------------------------------------------ gloo, world_size 2
def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "pytorch_nested-unet_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from dist import GradBuckets
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((7, 5), (33,), (64, 3, 3, 3), (5,), (1000,))]
    gb = GradBuckets(params, bucket_bytes=4096)
    gb.begin("cpu")
    for i, p in enumerate(params):
        gb.view(p).copy_(torch.full(p.shape, float((rank + 1) * (i + 1))))
        if i != 3:                       # parameter 3 never reports: finish() must flush its bucket
            gb.ready(p)
    gb.finish()
    ok = all(torch.allclose(gb.view(p), torch.full(p.shape, 1.5 * (i + 1))) for i, p in enumerate(params))
    # SyncBN exchange: all-reduced [sum, sumsq] + global count == statistics of the concatenated batch
    g = torch.Generator().manual_seed(5)
    full = torch.randn(8, 16, 6, 6, generator=g, dtype=torch.float64)
    mine = full[rank * 4:(rank + 1) * 4]
    sums = torch.stack([mine.sum((0, 2, 3)), (mine * mine).sum((0, 2, 3))])
    dist.all_reduce(sums)
    cnt = full.numel() / 16
    mean = sums[0] / cnt
    var = sums[1] / cnt - mean * mean
    ok = ok and torch.allclose(mean, full.mean((0, 2, 3))) and torch.allclose(var, full.var((0, 2, 3), unbiased=False))
    # IoU: counts are all-reduced, not ratios
    c = torch.tensor([10 * (rank + 1), 40], dtype=torch.int64)
    dist.all_reduce(c)
    ok = ok and c.tolist() == [30, 80]
    q.put((rank, bool(ok), len(gb.bucket_ranges)))
    dist.destroy_process_group()


def test_grad_buckets_and_syncbn_exchange_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] >= 2  # more than one bucket was exercised


def test_kernel_selection_policy_for_materialised_operands():
    """Pure host logic: which layers run on the stacked Cout = 32 kernel, and where the TMA width limit switches the engine back
    to the fused loaders."""
    import archs
    import engine
    m = archs.NestedUNet(1, 3, True)
    specs = engine._state(m).specs
    stacked = sorted((s.node, s.which, s.cin_pad) for s in specs
                     if engine._stacked(s, 4, 96 >> s.level, 96 >> s.level, engine._cin_pad(s, True)))
    assert stacked == [("0_1", 1, 96), ("0_2", 1, 128), ("0_3", 1, 160), ("0_4", 1, 192)]
    # 256-pixel rows exceed the 254-pixel tensor-load box: the kernels describe them as column strips, the same layers stay stacked
    assert sorted(s.node for s in specs if s.level == 0 and engine._stacked(s, 4, 256, 256, engine._cin_pad(s, True))) == ["0_1", "0_2", "0_3", "0_4"]
    assert engine._tma_width(254) and not engine._tma_width(255) and engine._mat_width(255) and engine._mat_width(512)
    stem = specs[0]
    assert engine._cin_pad(stem, True) == 32 and engine._cin_pad(stem, False) == 16
    u = archs.UNet(1, 3)
    us = engine._state(u).specs
    assert [(s.node, s.cin_pad) for s in us if engine._stacked(s, 2, 64 >> s.level, 64 >> s.level, engine._cin_pad(s, True))] == [("0_4", 96)]


def test_checkpoint_layout_interchanges_with_the_reference():
    """model.pth interchange (trains.py:344-346, val.py:58-59).  tests/golden/make_golden_r02.py recorded, with the UNMODIFIED reference
    modules: the (key, shape, dtype) manifest of the state_dict they save, and that a checkpoint they wrote loads (strict) in this
    repository's modules, is saved again here and loads (strict) back in the reference with identical values.  Here: every model
    this package builds has exactly that manifest, and a checkpoint survives torch.save / torch.load / load_state_dict bit for bit."""
    import io
    import archs
    import goldens
    rec = goldens.load("steps_r02.pt")["_checkpoints"]
    for name, (arch, k, ds) in (("NestedUNet_ds", ("NestedUNet", 1, True)), ("NestedUNet", ("NestedUNet", 1, False)), ("UNet", ("UNet", 1, False)),
                                ("NestedUNet_c4_ds", ("NestedUNet", 4, True))):
        assert rec[name]["round_trip_equal"]
        torch.manual_seed(1)
        m = getattr(archs, arch)(k, 3, ds)
        mine = [(key, tuple(v.shape), str(v.dtype)) for key, v in m.state_dict().items()]
        assert mine == [tuple(x) for x in rec[name]["manifest"]], name
        buf = io.BytesIO()
        torch.save(m.state_dict(), buf)
        buf.seek(0)
        torch.manual_seed(2)
        m2 = getattr(archs, arch)(k, 3, ds)
        m2.load_state_dict(torch.load(buf), strict=True)
        assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_loop_helpers_host_side(tmp_path):
    """utils.settle_gc (collect + freeze once), bench.ClockSampler's timestamp window, and the grouped criterion's bookkeeping
    (losses.mark_head_group / the NUNET_NO_HEAD_GROUP switch) — the parts of the loop around the step that need no GPU."""
    import datetime
    import gc
    import importlib.util
    import time
    import utils
    utils.settle_gc.done = False
    before = gc.get_freeze_count()
    utils.settle_gc()
    frozen = gc.get_freeze_count()
    assert frozen > before and utils.settle_gc.done
    utils.settle_gc()                                   # idempotent ...
    assert gc.get_freeze_count() == frozen
    utils.settle_gc(again=True)                         # ... unless asked again (bench.py: before each timed loop)
    assert gc.get_freeze_count() >= frozen
    gc.unfreeze()
    utils.settle_gc.done = False

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("nunet_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c = bench.ClockSampler(0, enabled=False)
    assert c.summary() == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    c.out = open(tmp_path / "smi.csv", "w+")
    now = datetime.datetime.now()

    def row(dt_ms, mhz, cap):
        ts = (now + datetime.timedelta(milliseconds=dt_ms)).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        return "%s, %d, 1965, Not Active, Not Active, Not Active, %s\n" % (ts, mhz, cap)
    c.out.write(row(-5000, 300, "Not Active"))          # before the timed window: ignored
    c.out.write(row(50, 1740, "Active") + row(250, 1800, "Active") + row(450, 1710, "Not Active"))
    c.out.write("garbage line\n")

    class _Proc(object):
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0
    c.proc = _Proc()
    c.t0 = time.time() - 0.01
    time.sleep(0.6)
    c.stop()
    assert c.summary() == {"sm_mhz": 1740, "sm_max_mhz": 1965, "reasons": ["sw_power_cap"], "samples": 3}

    import losses
    outs = [torch.zeros(2, 1, 4, 4) for _ in range(4)]
    marked = losses.mark_head_group(outs)
    assert marked is outs and all(o.__dict__["_nunet_head_group"][1] == k for k, o in enumerate(outs))
    assert all(o.__dict__["_nunet_head_group"][0] is outs[0].__dict__["_nunet_head_group"][0] for o in outs)
    os.environ["NUNET_NO_HEAD_GROUP"] = "1"
    try:
        plain = [torch.zeros(2, 1, 4, 4) for _ in range(4)]
        losses.mark_head_group(plain)
        assert all("_nunet_head_group" not in o.__dict__ for o in plain)
    finally:
        del os.environ["NUNET_NO_HEAD_GROUP"]


def _fake_loss_ops(monkeypatch):
    """torch-CPU stand-ins for the loss entry points of ops.py (losses.py:107-117 restated), so that the bookkeeping of the grouped
    criterion — which head is evaluated when, what is cached, what autograd receives, what keeps what alive — runs without a GPU.
    The kernels themselves are compared with the reference on the GPU (tests/test_head_loss_gpu.py)."""
    import ops
    calls = {"multi_fwd": 0, "multi_bwd": 0, "single_fwd": 0, "single_bwd": 0}

    def fwd(x, t):
        n = x.shape[0]
        p = torch.sigmoid(x).view(n, -1)
        tt = t.view(n, -1)
        sums = torch.stack([(p * tt).sum(1), p.sum(1), tt.sum(1)], 1).double()
        bce = torch.nn.functional.binary_cross_entropy_with_logits(x, t)
        dice = (2 * sums[:, 0] + 1e-5) / (sums[:, 1] + sums[:, 2] + 1e-5)
        loss = (0.5 * bce.double() + 1 - dice.sum() / n).float()
        o_, t_ = torch.sigmoid(x) > 0.5, t > 0.5
        return loss, sums, torch.stack([(o_ & t_).sum(), (o_ | t_).sum()]).long()

    def bwd(x, t, g, gmul):
        with torch.enable_grad():      # called from inside a custom Function's backward, where grad mode is off
            xr = x.detach().clone().requires_grad_(True)
            fwd_loss = fwd(xr, t)[0]
            (gx,) = torch.autograd.grad(fwd_loss, xr)
        return gx * (float(g) if g is not None else 1.0) * gmul

    def loss_iou_fwd(logits, target, want_counts=True):
        calls["single_fwd"] += 1
        return fwd(logits, target)

    def loss_bwd(logits, target, sums, gscale, gmul=1.0, dlogits=None):
        calls["single_bwd"] += 1
        return bwd(logits, target, gscale, gmul)

    def loss_iou_fwd_multi(logits_list, target):
        calls["multi_fwd"] += 1
        r = [fwd(x, target) for x in logits_list]
        return torch.stack([a for a, _, _ in r]), torch.stack([b for _, b, _ in r]), torch.stack([c for _, _, c in r])

    def loss_bwd_multi(logits_list, target, sums, gscales, gmul=1.0):
        calls["multi_bwd"] += 1
        return [bwd(x, target, g, gmul) if g is not None else None for x, g in zip(logits_list, gscales)]

    for name, fn in (("loss_iou_fwd", loss_iou_fwd), ("loss_bwd", loss_bwd), ("loss_iou_fwd_multi", loss_iou_fwd_multi),
                     ("loss_bwd_multi", loss_bwd_multi), ("require_cuda", lambda t, what: None), ("is_device_tensor", lambda t: True)):
        monkeypatch.setattr(ops, name, fn)
    return calls


def test_grouped_criterion_bookkeeping(monkeypatch):
    """losses.mark_head_group / BCEDiceLoss on the deep-supervision list (trains.py:120-123): one evaluation for the whole loop, the
    per-head path's values and gradients, dissolution after the last head, a partial loop, another target, no-grad mode, and — the
    point of the weak references — the logits die with their last reference, without the cyclic garbage collector."""
    import gc
    import weakref
    import losses
    calls = _fake_loss_ops(monkeypatch)
    crit = losses.BCEDiceLoss()
    g = torch.Generator().manual_seed(3)
    base = [torch.randn(3, 1, 8, 8, generator=g) for _ in range(4)]
    t = (torch.rand(3, 1, 8, 8, generator=g) > 0.5).float()
    t2 = (torch.rand(3, 1, 8, 8, generator=g) > 0.5).float()
    a = [b.clone().requires_grad_(True) for b in base]
    r = [b.clone().requires_grad_(True) for b in base]
    heads = losses.mark_head_group([x * 1.0 for x in a])
    group = heads[0].__dict__["_nunet_head_group"][0]
    la = [crit(h, t) for h in heads]
    assert calls["multi_fwd"] == 1 and calls["single_fwd"] == 0          # ONE evaluation served the whole loop
    assert group["tensors"] is None and group["losses"] is None          # ... and the group dissolved after the last head
    lr = [crit(x * 1.0, t) for x in r]
    assert calls["single_fwd"] == 4
    for k in range(4):
        assert torch.equal(la[k].detach(), lr[k].detach())
        assert heads[k].__dict__["_nunet_iou"][3].tolist() == losses.ops.loss_iou_fwd(base[k], t)[2].tolist()
    (sum(la) / 4).backward()
    (sum(lr) / 4).backward()
    assert calls["multi_bwd"] == 1
    for k in range(4):
        assert torch.allclose(a[k].grad, r[k].grad, rtol=0, atol=1e-7)
    # a fifth call on a head of a dissolved group takes the ordinary path
    n0 = calls["single_fwd"]
    again = crit(heads[2], t)
    assert calls["single_fwd"] == n0 + 1 and torch.equal(again.detach(), la[2].detach())

    # no cycle: with the cyclic collector off, the logits are freed when their last reference goes
    gc.collect()
    gc.disable()
    try:
        heads = losses.mark_head_group([x * 1.0 for x in a])
        refs = [weakref.ref(h) for h in heads]
        total = sum(crit(h, t) for h in heads) / 4
        total.backward()
        del heads, total
        assert all(rf() is None for rf in refs)
        # a partial loop (two of four heads) and then backward: still no cycle once backward has released the saved tensors
        for x in a:
            x.grad = None
        heads = losses.mark_head_group([x * 1.0 for x in a])
        refs = [weakref.ref(h) for h in heads]
        n_multi = calls["multi_fwd"]
        tot = crit(heads[0], t2) + crit(heads[3], t2)
        assert calls["multi_fwd"] == n_multi + 1 and heads[0].__dict__["_nunet_head_group"][0]["tensors"] is not None
        tot.backward()
        ref0 = base[0].clone().requires_grad_(True)
        crit(ref0 * 1.0, t2).backward()
        assert torch.allclose(a[0].grad, ref0.grad, rtol=0, atol=1e-7) and float(a[1].grad.abs().max()) == 0.0 and float(a[3].grad.abs().max()) > 0.0
        del heads, tot
        assert all(rf() is None for rf in refs)
    finally:
        gc.enable()

    # another target re-evaluates the group; no-grad mode goes through the direct multi-head call
    heads = losses.mark_head_group([x * 1.0 for x in a])
    n_multi = calls["multi_fwd"]
    l0, l0b = crit(heads[0], t), crit(heads[0], t2)
    assert calls["multi_fwd"] == n_multi + 2 and not torch.equal(l0.detach(), l0b.detach())
    with torch.no_grad():
        heads = losses.mark_head_group([b.clone() for b in base])
        assert torch.equal(crit(heads[2], t), lr[2].detach()) and crit(heads[2], t).grad_fn is None
    # a list that does not qualify (mixed shapes) falls back to the per-head path
    odd = losses.mark_head_group([torch.randn(3, 1, 8, 8), torch.randn(3, 1, 4, 4)])
    n0 = calls["single_fwd"]
    crit(odd[0], t)
    assert calls["single_fwd"] == n0 + 1


def test_bench_roofline_model_matches_the_survey():
    """bench.conv_flops_per_image / bench.network_roofline are the denominator of the headline `roofline.frac`: pin them to the
    survey's numbers (9.7008 GF/img forward and 29.0866 GF/img training for NestedUNet-DS at 96x96, cross-checked there
    against torch's flop counter on the reference; 8d: ~25.5 MB/img forward, ~91 MB/img training, 5.30 ms aggregate at batch 256)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("nunet_bench2", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    fwd, train = bench.conv_flops_per_image("NestedUNet", 96)
    heads = 4 * 2.0 * 96 * 96 * 32                       # four 1x1 heads (0.15 GF/img each forward in the survey's table, batch 256)
    assert abs((fwd + heads) / 1e9 - 9.7008) < 0.01
    assert abs((train + 3 * heads) / 1e9 - 29.0866) < 0.03
    # work scales with the pixel count; UNet is the encoder + the last anti-diagonal
    f256, _ = bench.conv_flops_per_image("NestedUNet", 256)
    assert abs(f256 / fwd - (256 / 96) ** 2) < 1e-9
    fu, tu = bench.conv_flops_per_image("UNet", 96)
    assert 0 < fu < fwd and 0 < tu < train
    pk = {"hbm": 6541.1, "tc_burst": 1658.4, "tc_sustained": 1390.8}
    net = bench.network_roofline("NestedUNet", 96, 256, True, pk)
    assert abs(net["flops"] - train * 256) < 1e-3 * train * 256
    assert 88e6 < net["bytes"] / 256 < 93e6                                  # survey: 91.0 MB per image-step
    assert 5.2 < net["ms"]["burst"] < 5.4 and 5.9 < net["ms"]["sustained"] < 6.2      # survey 8d: 5.30 ms at the burst peak
    assert 0.15 < net["hbm_bound_ms"]["sustained"] / net["ms"]["sustained"] < 0.35
    # more bandwidth can only lower it; a 1-class and a 4-class head differ by the fp32 logits traffic only
    fast = bench.network_roofline("NestedUNet", 96, 256, True, dict(pk, hbm=2 * pk["hbm"]))
    assert fast["ms"]["sustained"] < net["ms"]["sustained"]
    c4 = bench.network_roofline("NestedUNet", 96, 256, True, pk, num_classes=4)
    assert abs((c4["bytes"] - net["bytes"]) - 4 * 256 * 96 * 96 * 20 * 3) < 1.0
