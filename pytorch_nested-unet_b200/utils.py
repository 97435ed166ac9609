"""Drop-in for the reference's utils module (utils.py:4-33): str2bool, count_params, AverageMeter."""
import argparse


def str2bool(v):
    if v.lower() in ['true', 1]:
        return True
    elif v.lower() in ['false', 0]:
        return False
    else:
        raise argparse.ArgumentTypeError('Boolean value expected.')


def count_params(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


class AverageMeter(object):
    """Batch-size weighted running mean (utils.py:17-33)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def settle_gc(again=False):
    """One full garbage collection, then gc.freeze(): everything alive now (torch, the model, the captured graphs, the optimizer
    state) moves to the permanent generation and is never scanned again.  A training process has a large, static heap; a periodic
    generation-2 collection over it took 70 ms here and landed inside a step every ~40 steps — with the GPU queue only one step
    deep that is five idle steps, and under data parallelism every rank waits for the slowest one (measured: 2 GPUs, 20-step window,
    16.3 vs 14.4 ms/step).  Objects created afterwards are collected as usual.  Idempotent unless again=True."""
    import gc
    if getattr(settle_gc, "done", False) and not again:
        return
    gc.collect()
    gc.freeze()
    settle_gc.done = True


class DevicePrefetcher(object):
    """Wrapper of a loader of (tensor, ..., meta) host batches (train.train uses it): the host->device copy of batch i+1 is issued on
    a copy stream before batch i is handed out, so it overlaps batch i's step (the reference's loop copies synchronously on the
    compute stream, trains.py:113-114).  Pinned host tensors make the copy asynchronous; pageable ones still work.

    Steady state creates no CUDA resources (no cudaMalloc, no event / stream creation — each of those takes a driver-wide lock and
    stalls for as long as e.g. a concurrent `nvidia-smi` query holds it): the copy stream is cached per device, and batches land in a
    ring of `ring` device buffers per tensor position (re-created when a shape or dtype changes, e.g. the last short batch).  A slot is
    overwritten only after the work enqueued on the consumer's stream up to the moment the NEXT batch was requested has finished, so a
    batch is valid until `ring - 1` further batches have been requested — copy it if you need it longer.
    On its own the overlap measured no gain while the loop still drained the launch queue with `loss.item()` every step
    (tools/e2e_ab.py); together with StepScalars below the queue stays full and the copy is hidden."""

    _streams = {}

    def __init__(self, loader, device=None, ring=3):
        self.loader, self.device, self.ring = loader, device, max(2, int(ring))
        self._bufs = {}      # (position in the batch tuple, ring slot) -> device tensor
        self._ready = None   # per ring slot: copy finished (recorded on the copy stream)
        self._free = None    # per ring slot: consumer finished (recorded on the consumer's stream)
        self._used = None

    def __len__(self):
        return len(self.loader)

    def _copy_stream(self, dev):
        import torch
        key = (dev.type, dev.index)
        if key not in DevicePrefetcher._streams:
            DevicePrefetcher._streams[key] = torch.cuda.Stream(device=dev)
        return DevicePrefetcher._streams[key]

    def __iter__(self):
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if self.device is None else torch.device(self.device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        copy_stream = self._copy_stream(dev)
        if self._ready is None:
            self._ready = [torch.cuda.Event() for _ in range(self.ring)]
            self._free = [torch.cuda.Event() for _ in range(self.ring)]
            self._used = [False] * self.ring
        it = iter(self.loader)
        cur = torch.cuda.current_stream(dev)

        def load(n):
            try:
                batch = next(it)
            except StopIteration:
                return None
            k = n % self.ring
            out = []
            with torch.cuda.stream(copy_stream):
                if self._used[k]:
                    copy_stream.wait_event(self._free[k])          # the previous tenant of this slot has been consumed
                for j, b in enumerate(batch):
                    if torch.is_tensor(b) and not b.is_cuda:
                        buf = self._bufs.get((j, k))
                        if buf is None or buf.shape != b.shape or buf.dtype != b.dtype:
                            buf = torch.empty(b.shape, dtype=b.dtype, device=dev)
                            buf.record_stream(cur)                 # freed blocks must also wait for their readers' stream
                            self._bufs[(j, k)] = buf
                        buf.copy_(b, non_blocking=True)
                        out.append(buf)
                    else:
                        out.append(b)
                self._ready[k].record(copy_stream)
            return tuple(out), k

        n = 0
        nxt = load(0)
        while nxt is not None:
            out, k = nxt
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(self._ready[k])
            nxt = load(n + 1)
            try:
                yield out
            finally:                                                   # also when the consumer leaves its loop early
                self._free[k].record(torch.cuda.current_stream(dev))   # everything enqueued on batch n so far
                self._used[k] = True
            n += 1


class StepScalars(object):
    """The step's scalars (loss, IoU intersection / union counts) read WITHOUT draining the launch queue.

    The reference reads `loss.item()` and the IoU on the host after every step (trains.py:124,133-134).  A host read placed after the
    whole step stalls the CPU until the GPU has finished it, and the GPU then idles while the next step is being enqueued (copies,
    input kernels, graph launches).  Here submit() — called as soon as the loss and the counts exist, i.e. BEFORE backward is
    enqueued — issues their device->host copy into pinned memory behind the forward pass (asynchronous), and collect() — called after
    backward and the optimizer step have been enqueued — waits for that copy only: the CPU resumes when the forward pass is done and
    enqueues the next step while the GPU is still busy with backward.  Same values, same iteration, one copy per step; collect(keep=k)
    leaves the newest k submissions in flight for loops that want to run further ahead.
    Under data parallelism (reduce=True) the loss is averaged and the integer counts summed over the ranks on the device first
    (counts < 2^53 are exact in the fp64 exchange buffer)."""

    def __init__(self, group=None, reduce=False):
        self.group, self.reduce = group, reduce
        self.slots, self.free = [], []
        self.step_end = None      # recorded by collect(): everything of the step enqueued so far

    def before_backward(self):
        """Call right before `loss.backward()`: waits until the PREVIOUS step has left the GPU.  The forward pass of this step is
        already queued behind it, so the GPU keeps working while the host waits, and backward + optimizer are then enqueued while the
        GPU runs this step's forward.  Why: without it the captured backward graph is launched while its own previous launch is still
        executing, and `loss.backward()` then occasionally blocked the host for 50-190 ms (bench.py's host trace: `bwd` phase 51 / 163 /
        185 ms instead of 1.3 ms, GPU idle afterwards).  With it the stalls are rarer and smaller (one of 27 ms in 1 380 iterations) but
        not gone — their cause is open (profiles/r02_summary.md 6)."""
        if self.step_end is not None:
            self.step_end.synchronize()

    def _slot(self):
        import torch
        if self.free:
            return self.free.pop()
        return dict(loss=torch.empty(1, dtype=torch.float32).pin_memory(), counts=torch.empty(2, dtype=torch.int64).pin_memory(),
                    pack=torch.empty(3, dtype=torch.float64).pin_memory(), ev=torch.cuda.Event(), packed=False, world=1)

    def submit(self, loss, counts):
        """loss: 0-dim (or 1-element) CUDA tensor; counts: int64[2] CUDA tensor (metrics.iou_counts_device)."""
        import torch
        sl = self._slot()
        world = 1
        if self.reduce:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                world = dist.get_world_size(self.group)
        if world > 1:
            import torch.distributed as dist
            pk = torch.cat([loss.detach().reshape(1).double(), counts.double()])
            dist.all_reduce(pk, op=dist.ReduceOp.SUM, group=self.group)
            sl["pack"].copy_(pk, non_blocking=True)
            sl["packed"], sl["world"] = True, world
        else:
            sl["loss"].copy_(loss.detach().reshape(1), non_blocking=True)
            sl["counts"].copy_(counts, non_blocking=True)
            sl["packed"], sl["world"] = False, 1
        sl["ev"].record()
        self.slots.append(sl)

    def collect(self, keep=0):
        """(loss, intersection, union) of every submission except the newest `keep`, oldest first.  Call after the optimizer step has
        been enqueued: the end of the step is marked here for before_backward()."""
        import torch
        if self.step_end is None:
            self.step_end = torch.cuda.Event()
        self.step_end.record()
        out = []
        while len(self.slots) > keep:
            sl = self.slots.pop(0)
            sl["ev"].synchronize()
            if sl["packed"]:
                l, i, u = sl["pack"].tolist()
                out.append((l / sl["world"], int(round(i)), int(round(u))))
            else:
                out.append((float(sl["loss"][0]), int(sl["counts"][0]), int(sl["counts"][1])))
            self.free.append(sl)
        return out


def device_target(target, aug=None):
    """Loss / metric target from what the loader yields: the reference's fp32 NCHW mask tensor is passed through; a raw uint8 NHWC mask
    stack [N,H,W,K] (0 / 255, before dataset.py:73-74's transpose and /255) is converted on the GPU (ops.mask_u8)."""
    import torch
    if torch.is_tensor(target) and target.dtype == torch.uint8:
        import ops
        return ops.mask_u8(target.contiguous(), aug=aug)
    return target
