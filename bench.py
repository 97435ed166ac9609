#!/usr/bin/env python
"""bench.py — NestedUNet (deep supervision) train images/sec @96x96 bf16 on N B200s (BASELINE.json metric, config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--batch B] [--size S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one full iteration of the reference's training loop body (trains.py:113-136): forward, BCEDice loss (mean
over the four deep-supervision heads), iou_score on the last head, zero_grad, backward, SGD(momentum, weight decay)
step.  `value` times it with the batch already resident in HBM; `e2e` includes the pinned-host -> device copy of
every step's batch and the device -> host read of loss and IoU.  `--impl reference` times the reference's CPU path
(the oracle port of it; /root/reference does not exist on the GPU box) on the host cores.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pytorch_nested-unet_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "NestedUNet(deep_supervision) train images/sec @96x96"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch (weak scaling)")
    ap.add_argument("--size", type=int, default=96)
    ap.add_argument("--arch", default="NestedUNet")
    ap.add_argument("--no-ds", action="store_true")
    ap.add_argument("--classes", type=int, default=1, help="num_classes (BASELINE config 4: 4); the CPU legs run the 1-class reference step only")
    ap.add_argument("--no-sync-bn", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-layers", action="store_true", help="also write profiles/layers_latest.json")
    ap.add_argument("--e2e-loop", default="auto", choices=["auto", "overlap", "sync"],
                    help="what train.train() does: overlap (prefetcher + scalars read before backward is waited for) on one GPU, sync (the reference's "
                         "placement: copy on the compute stream, host reads after the step) under data parallelism; auto picks by world size")
    ap.add_argument("--kernels-only", action="store_true", help="profiler runs (ncu): only the warm-up and the timed device-resident steps")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "tc_burst": float(p["bf16_tflops"]), "tc_sustained": float(p["bf16_tflops_sustained"]),
                "src": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm": 6650.0, "tc_burst": 1590.0, "tc_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------ algorithmic work
def conv_flops_per_image(arch, size, input_channels=3):
    """2*M*Cout*9*Cin per conv; train = 3x forward except 2x for the stem conv (no dgrad). SURVEY.md 8d."""
    import engine
    graph = engine.nested_graph() if arch == "NestedUNet" else engine.unet_graph()
    nb = engine.NB_FILTER
    fwd = train = 0.0
    for node, srcs in graph:
        lvl = int(node[0])
        m = (size >> lvl) ** 2
        cout = nb[lvl]
        cin = 0
        for s, kind in srcs:
            cin += input_channels if kind == "input" else nb[int(s[0])]
        f1 = 2.0 * m * cout * 9 * cin
        f2 = 2.0 * m * cout * 9 * cout
        fwd += f1 + f2
        train += (2 * f1 if srcs[0][1] == "input" else 3 * f1) + 3 * f2
    return fwd, train


def network_roofline(arch, size, batch, ds, peaks_, num_classes=1, input_channels=3):
    """per-layer-max roofline of one training step, t = sum over the conv launches (forward, dgrad, wgrad) of
    max(FLOPs / P_tensor, algorithmic bytes / BW_hbm), plus the heads / loss as HBM traffic.  Byte model (bf16 activations): a conv
    reads each concatenated source at its STORED resolution (same-level M*C, pooled 4*M*C, upsampled M*C/4 elements), its weights, and
    writes one bf16 raw output; dgrad reads dy, y (ReLU mask / BatchNorm backward) and the weights and writes dx per source at stored
    resolution; wgrad reads the sources and dy and writes the fp32 dW.  Heads + loss per head and pixel: activation read twice (logits,
    dWh), logits / dlogits / target in fp32.  Reproduces the survey's 25.5 MB/img forward; 90.3 MB/img training (survey: 91.0)."""
    import engine
    graph = engine.nested_graph() if arch == "NestedUNet" else engine.unet_graph()
    nb = engine.NB_FILTER
    out = {"bytes": 0.0, "flops": 0.0, "hbm_bound_ms": {}, "ms": {}}
    tot = {k: 0.0 for k in ("burst", "sustained")}
    hbm_part = {k: 0.0 for k in ("burst", "sustained")}
    for node, srcs in graph:
        lvl = int(node[0])
        m = float(batch) * (size >> lvl) ** 2
        cout = nb[lvl]
        for which in (1, 2):
            src = [(kind, input_channels if kind == "input" else nb[int(s_[0])]) for s_, kind in srcs] if which == 1 else [("same", cout)]
            cin = sum(c for _, c in src)
            rd = sum({"input": 2 * m * c, "same": 2 * m * c, "pool": 8 * m * c, "up": 0.5 * m * c}[k] for k, c in src)
            wb = 2.0 * 9 * cin * cout
            fl = 2.0 * m * cout * 9 * cin
            stem = src[0][0] == "input"
            launches = [(fl, rd + 2 * m * cout + wb)]                                   # forward
            if not stem:
                launches.append((fl, 4 * m * cout + wb + rd))                           # dgrad
            launches.append((fl, rd + 2 * m * cout + 4.0 * 9 * cin * cout))             # wgrad
            for f, b in launches:
                out["flops"] += f
                out["bytes"] += b
                for key, ptc in (("burst", peaks_["tc_burst"]), ("sustained", peaks_["tc_sustained"])):
                    t_tc, t_hbm = f / (ptc * 1e12), b / (peaks_["hbm"] * 1e9)
                    tot[key] += max(t_tc, t_hbm)
                    if t_hbm >= t_tc:
                        hbm_part[key] += t_hbm
    heads = 4 if (ds and arch == "NestedUNet") else 1
    hb = heads * float(batch) * size * size * (2 * 64 + 20 * num_classes)
    out["bytes"] += hb
    for key in tot:
        tot[key] += hb / (peaks_["hbm"] * 1e9)
        hbm_part[key] += hb / (peaks_["hbm"] * 1e9)
        out["ms"][key] = tot[key] * 1e3
        out["hbm_bound_ms"][key] = hbm_part[key] * 1e3
    return out


class ClockSampler(object):
    """SM clock and throttle reasons DURING the timed regions (B200_PROFILING.md's clocks line): ONE `nvidia-smi -lms 200` process,
    started by rank 0 only and before the warm-up, so that its start-up (NVML attaches to every GPU of the box under a driver-wide
    lock) is over when the timing starts; its rows are filtered to the timed window by their own timestamps.  (Spawning a fresh
    nvidia-smi every 200 ms from every rank, as this bench first did, cost the 2-GPU run an occasional 25 ms stall: 14.5 vs 15.8
    ms/step between otherwise identical runs.)"""

    FIELDS = "timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, enabled=True):
        self.proc, self.t0, self.t1, self.rows = None, None, None, []
        if enabled:
            try:
                import tempfile
                self.out = tempfile.TemporaryFile(mode="w+")
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "200"],
                                             stdout=self.out, stderr=subprocess.DEVNULL)
            except Exception:
                self.proc = None

    def mark_start(self):
        self.t0 = time.time()

    def stop(self):
        self.t1 = time.time()
        if self.proc is None:
            return
        out = ""
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
        try:
            self.out.seek(0)
            out = self.out.read()
            self.out.close()
        except Exception:
            pass
        self.proc = None
        import datetime
        for line in out.splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                continue
            if self.t0 is not None and self.t0 <= ts <= self.t1:
                self.rows.append(c[1:])

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_step_rate(arch, ds, n, size, steps, warmup):
    """Times the oracle port of the reference's train step (trains.py:113-136) on the host cores."""
    from oracle import nunet_oracle as O
    import archs
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = getattr(archs, arch)(1, 3, ds)          # parameter container only (CPU): same init as the reference
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x, t = O.synthetic_batch(n, size, size, 1, 3, 1, "dsb")
    bufs = {}
    times = []
    first = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        sd, loss, iou, _, outs = O.train_step(sd, x, t, arch, ds, bufs=bufs)
        dt = time.perf_counter() - t0
        if first is None:
            first = (x, outs if isinstance(outs, list) else [outs], loss)     # step-0 logits at default init: the numerics check of the native arm
        if i >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return n / med, med, torch.get_num_threads(), first


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 16
    steps = max(1, min(args.steps, 5))
    warm = 1
    ds = not args.no_ds
    rate, med, cores, _ = cpu_reference_step_rate(args.arch, ds, n, args.size, steps, warm)
    sample = "oracle port of trains.py:113-136, %s ds=%s, batch %d @%dx%d fp32 on CPU, median of %d steps after %d warm-up" % (
        args.arch, ds, n, args.size, args.size, steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
        "data": "synthetic",
        "config": {"workload": "%s(num_classes=1, input_channels=3, deep_supervision=%s) %dx%d, batch %d per GPU, full train step "
                               "(fwd + BCEDice + IoU + bwd + SGD)" % (args.arch, ds, args.size, args.size, args.batch),
                   "global_batch": args.batch * args.gpus, "parallelism": "cpu (reference step on the host cores)",
                   "sample": "each step is a batch-%d sample of the workload (a batch-%d CPU step takes ~10 s)" % (n, args.batch)},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ native arm
def run_native(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import archs
    import losses
    import metrics
    import ops
    ds = not args.no_ds
    torch.manual_seed(0)
    model = getattr(archs, args.arch)(args.classes, 3, ds).to(dev)
    crit = losses.BCEDiceLoss().to(dev)
    if world > 1:
        import dist as nunet_dist
        model = nunet_dist.DataParallel(model, sync_bn=not args.no_sync_bn)
    import optim as nunet_optim
    opt = nunet_optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)     # trains.py:229-231 defaults, fused step (optim.py)
    B, S = args.batch, args.size
    # Synthetic batch in the form the reference's loader reads it from disk: uint8 HWC images and 0/255 masks (dataset.py:58-68).  The
    # loader's per-pixel arithmetic — albumentations Normalize() (trains.py:266), the extra /255 (dataset.py:71), mask/255 and the
    # HWC->CHW transposes (dataset.py:72-74) — runs on the GPU (ops.input_u8 / ops.mask_u8), so a step ships 4x fewer bytes.
    g = torch.Generator().manual_seed(1 + rank)
    x_host = torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8).pin_memory()
    t_host = ((torch.rand(B, S, S, args.classes, generator=g) > 0.5).to(torch.uint8) * 255).pin_memory()
    x_dev, t_dev = x_host.to(dev), ops.mask_u8(t_host.to(dev))

    STAMPS = None

    def step(x, t, scalars=None):
        if STAMPS is not None:
            STAMPS.append(time.perf_counter())
        outputs = model(x)
        if isinstance(outputs, list):
            loss = 0
            for o in outputs:
                loss = loss + crit(o, t)
            loss = loss / len(outputs)
            last = outputs[-1]
        else:
            loss = crit(outputs, t)
            last = outputs
        counts = metrics_counts(last, t)
        if scalars is not None:
            scalars.submit(loss, counts)            # device->host copy of the step's scalars, issued before backward (utils.StepScalars)
            if not os.environ.get("NUNET_NO_BWD_FENCE"):
                scalars.before_backward()           # the previous step has left the GPU before this backward graph is launched again
        if STAMPS is not None:
            STAMPS.append(time.perf_counter())
        opt.zero_grad()
        loss.backward()
        if STAMPS is not None:
            STAMPS.append(time.perf_counter())
        opt.step()
        if STAMPS is not None:
            STAMPS.append(time.perf_counter())
        return loss, counts

    def metrics_counts(last, t):
        # device-side integer counts (metrics.iou_score without the host read, which the e2e loop does explicitly)
        cached = last.__dict__.get("_nunet_iou")
        return cached[3] if cached is not None else ops.iou_counts(last.detach(), t)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    model.train()
    # at least three untimed steps: the first runs eagerly, the second captures the CUDA graphs of the step, the third replays
    warm_run = max(args.warmup, 3)
    sampler = ClockSampler(local, enabled=(rank == 0 and not os.environ.get("NUNET_NO_CLOCKS")))      # started before the warm-up, see the class
    for _ in range(warm_run):
        step(x_dev, t_dev)
    import utils
    if not os.environ.get("NUNET_NO_GC_FREEZE"):
        utils.settle_gc()            # what train.train() does after its third step (a 70 ms generation-2 collection otherwise lands in a timed step)
    barrier()
    sampler.mark_start()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, counts = step(x_dev, t_dev)
    e1.record()
    barrier()
    launches = ops.launch_count()
    ms = e0.elapsed_time(e1)
    if args.kernels_only:
        sampler.stop()
        if rank == 0:
            print(json.dumps({"kernels_only": True, "ms_per_step": ms / args.steps, "gpu_launches": launches}))
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- end-to-end through the public API, exactly train.py's loop (train.train): every step's batch comes from pinned host
    #      memory through utils.DevicePrefetcher (the copy of batch i+1 is issued on a copy stream when batch i is handed out), and the
    #      step's loss and IoU counts are copied to the host every step and read there in the same iteration through
    #      utils.StepScalars (copy issued before backward is enqueued, waited for after the optimizer step is enqueued).

    class HostBatches(object):            # the loader: `n` pinned host batches per pass
        n = 20                            # untimed pass (ring buffers, pinned result slots, events; the rare host stalls of the first second)

        def __iter__(self):
            for _ in range(self.n):
                yield x_host, t_host, None

    batches = HostBatches()
    loader = utils.DevicePrefetcher(batches, dev)
    scal = utils.StepScalars()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if not os.environ.get("NUNET_NO_GC_FREEZE"):
        utils.settle_gc(again=True)
    e2e_loop = args.e2e_loop if args.e2e_loop != "auto" else ("overlap" if world == 1 else "sync")
    if e2e_loop == "sync":                # the reference's own placement (copy on the compute stream, host reads after the step): train.train() under DP
        for _ in range(2):
            step(x_host.to(dev, non_blocking=True), ops.mask_u8(t_host.to(dev, non_blocking=True)))
        barrier()
        f0.record()
        for _ in range(args.steps):
            loss, counts = step(x_host.to(dev, non_blocking=True), ops.mask_u8(t_host.to(dev, non_blocking=True)))
            lv = loss.item()
            i_, u_ = counts.tolist()
            iou = (i_ + 1e-5) / (u_ + 1e-5)
        f1.record()
    else:
        for xb, tb, _ in loader:          # untimed: the ring buffers, pinned result slots and events are created here
            step(xb, ops.mask_u8(tb), scal)
            scal.collect()
        barrier()
        batches.n = args.steps
        import gc
        gc_t = {}

        def gc_cb(phase, info):      # a collection that takes more than 1 ms inside the timed loop is reported on stderr
            if phase == "start":
                gc_t["t"] = time.perf_counter()
            else:
                dt = (time.perf_counter() - gc_t.get("t", time.perf_counter())) * 1e3
                if dt > 1.0:
                    print("rank %d gc gen%d %.1f ms collected %d" % (rank, info["generation"], dt, info["collected"]), file=sys.stderr)
        gc.callbacks.append(gc_cb)
        f0.record()
        trace = True if os.environ.get("NUNET_E2E_TRACE") else None       # also print the phase medians when nothing was slow
        STAMPS = []
        iter_t = [time.perf_counter()]
        for xb, tb, _ in loader:
            loss, counts = step(xb, ops.mask_u8(tb), scal)
            (lv, i_, u_), = scal.collect()
            if STAMPS is not None:
                STAMPS.append(time.perf_counter())
            iter_t.append(time.perf_counter())
            iou = (i_ + 1e-5) / (u_ + 1e-5)
        f1.record()
        if len(iter_t) > 2:    # always: the three slowest host iterations (a stall shows up here with its position)
            dt = sorted(((iter_t[k + 1] - iter_t[k]) * 1e3, k + 1) for k in range(len(iter_t) - 1))
            print("rank %d e2e iterations: median %.2f ms, slowest %s" % (rank, dt[len(dt) // 2][0], [(k, round(v, 1)) for v, k in dt[-3:]]), file=sys.stderr)
        if STAMPS:   # per iteration (ms): loader + forward + loss + submit | backward | optimizer | wait for the scalars
            rows = []
            for k in range(len(STAMPS) // 5):
                a = STAMPS[5 * k:5 * k + 5]
                prev = STAMPS[5 * k - 1] if k else a[0]
                rows.append((round((a[0] - prev) * 1e3, 2), round((a[1] - a[0]) * 1e3, 2), round((a[2] - a[1]) * 1e3, 2), round((a[3] - a[2]) * 1e3, 2), round((a[4] - a[3]) * 1e3, 2)))
            tot = [sum(r) for r in rows]
            med = sorted(tot)[len(tot) // 2]
            slow = [(k, rows[k]) for k in range(len(rows)) if tot[k] > 1.15 * med]
            if slow or trace is not None:
                print("rank %d e2e host trace: median iteration %.2f ms; median phases (loader, fwd+loss, bwd, opt, wait) %s; iterations over 1.15x median: %s" % (
                    rank, med, [sorted(r[j] for r in rows)[len(rows) // 2] for j in range(5)], slow), file=sys.stderr)
            STAMPS = None
    barrier()
    sampler.stop()
    ms_e2e = f0.elapsed_time(f1)
    tm = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms, ms_e2e = tm.tolist()
    imgs = B * world * args.steps
    value = imgs / (ms * 1e-3)
    e2e = imgs / (ms_e2e * 1e-3)
    # ---- roofline of the dominant kernel class (implicit-GEMM conv: forward + dgrad + wgrad launches), timed per launch
    pk = peaks()
    raw = model.module if hasattr(model, "module") else model
    roof = None
    layer_rows = []
    # every rank runs the extra (untimed) profiling step: under DP it contains collectives
    import engine
    side_saved, lvl_saved = engine.WGRAD_SIDE_STREAM, engine.LEVEL_STREAMS
    engine.WGRAD_SIDE_STREAM = False            # per-kernel CUDA events need every launch on the one (launching) stream
    engine.LEVEL_STREAMS = False
    saved_graph = engine.GRAPH_TRAIN
    engine.GRAPH_TRAIN = False
    step(x_dev, t_dev)                          # untimed eager single-stream step: warms this schedule's allocator pool
    torch.cuda.synchronize()
    ops.enable_kernel_timing(rank == 0)
    step(x_dev, t_dev)
    torch.cuda.synchronize()
    engine.GRAPH_TRAIN = saved_graph
    timing = ops.collect_kernel_timing() if rank == 0 else None
    ops.enable_kernel_timing(False)
    engine.WGRAD_SIDE_STREAM, engine.LEVEL_STREAMS = side_saved, lvl_saved
    barrier()
    if rank == 0:
        conv_ms = sum(v for k, v in timing["by_class"].items() if k in ("conv_fwd", "conv_dgrad", "conv_wgrad"))
        total_ms = sum(timing["by_class"].values())
        _, train_fl = conv_flops_per_image(args.arch, S)
        tflops = train_fl * B / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        net = network_roofline(args.arch, S, B, ds, pk, num_classes=args.classes)
        step_ms = ms / args.steps
        # the dominant kernel: the generic tcgen05 implicit-GEMM conv (forward of the non-stacked layers + every dgrad)
        dom = [r for r in timing["launches"] if r["class"] in ("conv_fwd", "conv_dgrad") and r["meta"] and r["meta"][-1] == "igemm"]
        dom_ms = sum(r["ms"] for r in dom)
        dom_fl = sum(2.0 * r["meta"][0] * r["meta"][1] * r["meta"][2] * r["meta"][3] * r["meta"][4] * 9 for r in dom)
        dom_by = sum(2.0 * r["meta"][0] * r["meta"][1] * r["meta"][2] * (r["meta"][3] + r["meta"][4]) for r in dom)
        traffic, tsrc = None, None
        try:  # DRAM bytes per step from the committed ncu pass of this command (batch 256 @96x96 DS only)
            tfile = os.path.join(ROOT, "profiles", "r02_traffic.json")
            with open(tfile) as f:
                tj = json.load(f)
            if tj.get("batch") == B and tj.get("size") == S and args.arch == "NestedUNet" and ds:
                traffic, tsrc = tj["all_dram_bytes_per_step"], "profiles/r02_traffic.json (ncu dram__bytes_read+write over one step of this command; not re-measured in this run)"
        except Exception:
            pass
        hbm_share = net["hbm_bound_ms"]["sustained"] / net["ms"]["sustained"]
        roof = {
            # SURVEY 8d aggregate: sum over layers of max(FLOPs/P_tensor, bytes/BW_hbm) divided by the measured step
            "bound": "hbm" if hbm_share >= 0.5 else "tensor",
            "frac": net["ms"]["sustained"] / step_ms,
            "frac_burst_peaks": net["ms"]["burst"] / step_ms,
            # achieved / peak / unit follow `bound`: whole-step algorithmic FLOPs (or bytes) over the step time against that peak
            "achieved": (net["bytes"] / (step_ms * 1e-3) / 1e9) if hbm_share >= 0.5 else (train_fl * B / (step_ms * 1e-3) / 1e12),
            "peak": pk["hbm"] if hbm_share >= 0.5 else pk["tc_sustained"], "unit": "GB/s" if hbm_share >= 0.5 else "TFLOP/s",
            "achieved_hbm_gbs_algorithmic": net["bytes"] / (step_ms * 1e-3) / 1e9,
            "frac_note": "frac = per-layer-max network roofline (%.2f ms with the sustained bf16 peak, %.2f ms with the burst peak; %.0f %% of it in "
                         "HBM-bound layers) / measured ms_per_step — SURVEY 8d's aggregate, which is above achieved/peak because the HBM-bound "
                         "layers cannot run at the tensor peak; achieved = algorithmic FLOPs of the step / step time" % (
                             net["ms"]["sustained"], net["ms"]["burst"], 100 * hbm_share),
            "network_roofline_ms": net["ms"], "algorithmic_bytes_per_step": net["bytes"], "algorithmic_flops_per_step": train_fl * B,
            "traffic": traffic, "traffic_source": tsrc,
            "peak_source": pk["src"],
            "dominant_kernel": {"name": "conv3x3_igemm_kernel (tcgen05 implicit GEMM: forward of the non-stacked layers + every dgrad)",
                                "launches": len(dom), "ms_per_step": dom_ms, "share_of_kernel_time": dom_ms / total_ms if total_ms else None,
                                "achieved_tflops": dom_fl / (dom_ms * 1e-3) / 1e12 if dom_ms else None,
                                "frac_of_tensor_sustained": dom_fl / (dom_ms * 1e-3) / 1e12 / pk["tc_sustained"] if dom_ms else None,
                                "achieved_gbs_algorithmic": dom_by / (dom_ms * 1e-3) / 1e9 if dom_ms else None},
            "conv_class": {"ms_per_step": conv_ms, "achieved_tflops": tflops, "frac_of_tensor_sustained": tflops / pk["tc_sustained"]},
            "all_kernels_ms_per_step": total_ms,
            "class_ms": {k: round(v, 4) for k, v in sorted(timing["by_class"].items())}}
        layer_rows = timing["launches"]
        if args.profile_layers:
            os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
            with open(os.path.join(ROOT, "profiles", "layers_latest.json"), "w") as f:
                json.dump({"batch": B, "size": S, "arch": args.arch, "launches": layer_rows}, f, indent=1)
    cpu = None
    numerics = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.classes == 1:
        rate, med, cores, first = cpu_reference_step_rate(args.arch, ds, 16, S, 3, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "oracle port of the reference train step, %s ds=%s batch 16 @%dx%d fp32, median of 3 steps (%.2f s/step)" % (args.arch, ds, S, S, med)}
        # numerics of this arm on that very batch and initialisation (train mode, default init, the loader's input distribution)
        torch.manual_seed(0)
        m0 = getattr(archs, args.arch)(1, 3, ds).to(dev).train()
        with torch.no_grad():
            o0 = m0(first[0].to(dev))
        o0 = o0 if isinstance(o0, list) else [o0]
        numerics = {"train_mode_logits_relmax_per_head": [float((a_.cpu() - r_).abs().max() / r_.abs().max()) for a_, r_ in zip(o0, first[1])],
                    "train_mode_logits_rel_l2_per_head": [float((a_.cpu() - r_).norm() / r_.norm()) for a_, r_ in zip(o0, first[1])],
                    "against": "fp32 CPU oracle port, batch 16, default init (north_star: <= 2e-2 in bf16; an ideal bf16-operand pipeline gives "
                               "1e-2 .. 4e-2 from head 1 to head 4 in train mode at default init, SURVEY.md 8c finding 2)"}
        del m0
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "warmup_run": warm_run,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "%s(num_classes=%d, input_channels=3, deep_supervision=%s) %dx%d, batch %d per GPU, full train step "
                                   "(fwd + BCEDice + IoU + bwd + SGD)" % (args.arch, args.classes, ds, S, S, B),
                       "global_batch": B * world, "parallelism": "dp%d%s" % (world, "" if world == 1 else ("+syncbn" if not args.no_sync_bn else "")),
                       "l2": "working set per step (>2 GB of activations) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * x_host.element_size() + t_host.numel() * t_host.element_size()) * world,
                    "d2h_bytes_per_step": (4 + 16) * world, "ms_per_step": ms_e2e / args.steps,
                    "loop": e2e_loop,
                    "how": ("train.train()'s loop: pinned uint8 batch -> device every step (copy stream, one batch ahead), loss + IoU counts -> host "
                            "every step and read in the same iteration (copy issued before backward is enqueued)") if e2e_loop == "overlap" else
                           ("train.train()'s loop under data parallelism: pinned uint8 batch -> device on the compute stream every step, loss + IoU "
                            "counts read on the host after the step (the reference's placement, trains.py:113-134)")},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": roof,
            "cpu_baseline": cpu,
            "numerics": numerics,
            "last_loss": float(lv), "last_iou": float(iou),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # the step runs on ten CUDA streams (five levels + their weight-gradient streams): give each its own hardware queue, so that the
    # peer-waiting SyncBN tails of one stream can never sit in front of another stream's kernels (must be set before CUDA initialises)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    # the contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner) are sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
