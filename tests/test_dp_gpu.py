"""Data parallelism, one process per GPU.

  * 2 GPUs (NCCL, `gpurun --gpus 2`): the in-kernel SyncBN exchange (csrc/bn_tail.cuh, dist.PeerExchange) is held TIGHTLY — a conv
    with its BatchNorm tail and a gradient router with its BatchNorm-backward tail, run by two ranks on the halves of a batch, must
    give every rank the coefficients a single GPU computes on the whole batch (scale / shift / mean / invstd / running statistics,
    dgamma / dbeta / k1..k3 with the 1/world gradient scale) to fp32 round-off.  Then the whole DataParallel step (SyncBN in the
    kernels, captured graphs, flat-buffer all-reduce) against the single-process step on the concatenated batch.
  * 1 GPU (the driver's box): two ranks share cuda:0 over gloo (no in-kernel waiting between ranks on one device) and run the
    DataParallel step with SyncBN through all-reduced sums: same comparison.
At model level bf16 chaos limits how tight the comparison can be (a 1-ulp difference of a BatchNorm coefficient flips single bf16
roundings, which 30 ReLU layers amplify — tests/test_backward_consistency_gpu.py); the bound used is 3x what the single-process step
itself moves when its batch is merely permuted (same mathematics, different summation order)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup_paths():
    for p in (ROOT, os.path.join(ROOT, "pytorch_nested-unet_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)


def _batch():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 3, 64, 64, generator=g)
    t = (torch.rand(4, 1, 64, 64, generator=g) > 0.5).float()
    return x, t


def _step(model, x, t):
    import losses
    import metrics
    crit = losses.BCEDiceLoss()
    out = model(x)
    loss = sum(crit(o, t) for o in out) / len(out)
    counts = metrics.iou_counts(out[-1], t)
    loss.backward()
    return out, loss, counts


def _kernel_case(rank, world, dev):
    """conv + BN tail and router + BN-backward tail with the peer exchange, on this rank's half of a fixed batch."""
    import _lib as L
    import dist as nd
    import ops
    g = torch.Generator().manual_seed(21)
    N, H, W, cin, cout = 4, 16, 16, 32, 64
    a = torch.rand(N, H, W, cin, generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, generator=g) * 0.05
    gamma, beta, bias = (torch.randn(cout, generator=g) for _ in range(3))
    dx = torch.randn(N, H, W, cout, generator=g).to(torch.bfloat16)
    sl = slice(rank * N // world, (rank + 1) * N // world) if world > 1 else slice(0, N)
    n_loc = N // world if world > 1 else N
    a_, dx_ = a[sl].contiguous().to(dev), dx[sl].contiguous().to(dev)
    gamma, beta, bias, w = gamma.to(dev), beta.to(dev), bias.to(dev), w.to(dev)
    plan = ops.conv_plan(n_loc, H, W, cin, cout)
    wp = ops.pack_weights(w, L.PACK_FWD, plan)
    xch = nd.PeerExchange(2, dev) if world > 1 else None
    f32 = lambda: torch.empty(cout, device=dev)
    res = {}
    for rep in range(3):        # several epochs through the same slots (parity double buffering, epoch counters)
        rm, rv = torch.zeros(cout, device=dev), torch.ones(cout, device=dev)
        sc, sh, mu, rs = f32(), f32(), f32(), f32()
        counter = torch.zeros(2, dtype=torch.int32, device=dev)
        tail = ops.bn_tail(counter[0:1], float(N * H * W), gamma, beta, bias, rm, rv, 0.1, 1e-5, sc, sh, mu, rs, xch=xch.slot(0) if xch else None)
        y, _ = ops.conv3x3(n_loc, H, W, [ops.Source(L.SRC_ACT, a_, cin)], cout, wp, want_stats=True, tail=tail)
        dg, db, k1, k2, k3 = f32(), f32(), f32(), f32(), f32()
        bt = ops.bn_bwd_tail(counter[1:2], float(N * H * W), gamma, mu, rs, dg, db, k1, k2, k3, grad_scale=1.0 / world, xch=xch.slot(1) if xch else None)
        ops.grad_router(n_loc, H, W, cout, y, sc, sh, mu, rs, [ops.Consumer(dx_, cout, 0, L.SRC_SAME)], tail=bt)
        torch.cuda.synchronize(dev)
        res = {k: v.cpu() for k, v in dict(scale=sc, shift=sh, mean=mu, invstd=rs, rm=rm, rv=rv, dgamma=dg * world, dbeta=db * world, k1=k1, k2=k2, k3=k3).items()}
    return res


def _worker(rank, world, port, q, backend, same_device):
    _setup_paths()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    devi = 0 if same_device else rank
    torch.cuda.set_device(devi)
    dev = torch.device("cuda", devi)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import archs
    import dist as nd
    res = {"rank": rank}
    if backend == "nccl":
        res["kernel"] = _kernel_case(rank, world, dev)
    torch.manual_seed(0)
    model = nd.DataParallel(archs.NestedUNet(1, 3, True).to(dev), sync_bn=True, bucket_bytes=1 << 20)
    res["peer_syncbn"] = model.exchange is not None
    x, t = _batch()
    xs, ts = x[rank * 2:(rank + 1) * 2].to(dev), t[rank * 2:(rank + 1) * 2].to(dev)
    for it in range(3):          # eager step, capturing step, replayed step (under NCCL with the peer exchange; eager over gloo)
        model.module.zero_grad()
        torch.manual_seed(0)
        if it > 0:               # same weights and running statistics for every repetition
            model.module.load_state_dict(sd0)
        else:
            sd0 = {k: v.clone() for k, v in model.module.state_dict().items()}
        out, loss, counts = _step(model, xs, ts)
    import engine
    st = engine._state(model.module)
    res["captured"] = sum(1 for tg in st.__dict__.get("train_graphs", {}).values() if tg.fwd is not None and tg.bwd is not None)
    gl = nd.global_mean(float(loss.detach()), xs.device)
    giou = nd.global_iou_score(out[-1], ts)
    import metrics
    import utils
    scal = utils.StepScalars(reduce=True)        # train.train()'s read of the step's scalars under data parallelism
    scal.submit(loss.detach(), metrics.iou_counts_device(out[-1], ts))
    (sl, si, su), = scal.collect()
    torch.cuda.synchronize()
    res.update({"loss": gl, "iou": giou, "scalars": [sl, (si + 1e-5) / (su + 1e-5)], "logits": out[-1].detach().cpu(),
                "grads": {k: p.grad.detach().cpu() for k, p in model.module.named_parameters()},
                "rm": model.module.conv2_0.bn1.running_mean.cpu(), "rv": model.module.conv0_4.bn2.running_var.cpu()})
    q.put(_to_numpy(res))        # by value: torch tensors travel as shared-memory handles that die with this process
    dist.barrier()
    dist.destroy_process_group()


def _to_numpy(o):
    if torch.is_tensor(o):
        return ("__t__", o.numpy())
    if isinstance(o, dict):
        return {k: _to_numpy(v) for k, v in o.items()}
    return o


def _from_numpy(o):
    if isinstance(o, tuple) and len(o) == 2 and o[0] == "__t__":
        return torch.from_numpy(o[1])
    if isinstance(o, dict):
        return {k: _from_numpy(v) for k, v in o.items()}
    return o


def _run_ranks(backend, same_device):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, backend, same_device)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([_from_numpy(q.get(timeout=600)) for _ in procs], key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=120)
    return res


def _single(perm=None):
    import archs
    torch.manual_seed(0)
    model = archs.NestedUNet(1, 3, True).cuda()
    x, t = _batch()
    if perm is not None:
        x, t = x[perm], t[perm]
    out, loss, counts = _step(model, x.cuda(), t.cuda())
    torch.cuda.synchronize()
    return model, out, loss, counts


def _compare(res):
    model, out, loss, counts = _single()
    # what the same mathematics moves when only the summation order changes: the single-process step on the permuted batch
    perm = [2, 3, 0, 1]
    model_p, out_p, loss_p, _ = _single(perm)
    named, named_p = dict(model.named_parameters()), dict(model_p.named_parameters())
    inv = [perm.index(i) for i in range(4)]
    noise_logits = float((out_p[-1].detach()[inv] - out[-1].detach()).abs().max() / out[-1].abs().max())
    assert abs(res[0]["loss"] - float(loss.detach())) < max(2e-5, 3 * abs(float(loss_p.detach()) - float(loss.detach()))) and abs(res[1]["loss"] - res[0]["loss"]) < 1e-9
    I, U = counts
    for r in res:      # utils.StepScalars(reduce=True): global mean loss and IoU over the summed counts, same on every rank
        assert abs(r["scalars"][0] - r["loss"]) < 1e-6 and r["scalars"][1] == r["iou"]
    assert res[0]["iou"] == res[1]["iou"]
    assert abs(res[0]["iou"] - (I + 1e-5) / (U + 1e-5)) < 2e-3
    both = torch.cat([res[0]["logits"], res[1]["logits"]])
    dl = float((both - out[-1].detach().cpu()).abs().max() / out[-1].abs().max())
    assert dl <= max(1e-3, 3 * noise_logits), (dl, noise_logits)
    worst = (0.0, None, 0.0)
    for k, g0 in res[0]["grads"].items():
        assert torch.equal(g0, res[1]["grads"][k]), k                      # all ranks hold the same averaged gradient
        r = named[k].grad.cpu()
        if float(r.norm()) < 1e-9:
            continue
        noise = float((named_p[k].grad.cpu() - r).norm() / r.norm())
        rel = float((g0 - r).norm() / r.norm())
        if rel - 3 * noise > worst[0]:
            worst = (rel - 3 * noise, k, noise)
        assert rel <= max(2e-3, 3 * noise), (k, rel, noise)
    assert torch.allclose(res[0]["rm"], model.conv2_0.bn1.running_mean.cpu(), rtol=1e-4, atol=1e-6)
    assert torch.allclose(res[0]["rv"], model.conv0_4.bn2.running_var.cpu(), rtol=1e-4, atol=1e-6)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_dp2_nccl_peer_syncbn_and_captured_step():
    _setup_paths()
    res = _run_ranks("nccl", same_device=False)
    # kernel level, tight: both ranks hold the single-GPU coefficients of the whole batch
    ref = _kernel_case(0, 1, torch.device("cuda", 0))
    for r in res:
        assert r["peer_syncbn"], "the peer-memory exchange was not available on this box"
        for k, v in ref.items():
            tol = 2e-5 * float(v.abs().max()) + 1e-7
            assert float((r["kernel"][k] - v).abs().max()) <= tol, (k, float((r["kernel"][k] - v).abs().max()), tol)
        assert r["captured"] >= 1, "the data-parallel step was not captured"
    for k in ref:
        assert torch.equal(res[0]["kernel"][k], res[1]["kernel"][k]), k      # summed in rank order: identical bits on every rank
    _compare(res)


def test_dp2_on_one_gpu_over_gloo():
    """Runs on the driver's single-GPU box: two ranks on cuda:0, SyncBN sums and gradient buckets all-reduced over gloo."""
    _setup_paths()
    _compare(_run_ranks("gloo", same_device=True))
