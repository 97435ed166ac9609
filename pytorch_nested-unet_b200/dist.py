"""Data-parallel training of the fused engine: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The reference is single-GPU; this is the north_star's addition.  Exchange steps:
  1. gradient all-reduce (AVG) — the engine writes every parameter gradient straight into one flat fp32 buffer laid out
     in the order backward produces them; buckets of that buffer are all-reduced asynchronously as soon as their last
     gradient has been written, so the transfers overlap the remaining backward kernels (no copy-in / copy-out);
  2. SyncBN — per BatchNorm layer the fp64 [sum, sum of squares] (forward) and [sum g_z, sum g_z*yhat] (backward) are
     all-reduced and finalised with the GLOBAL count, which reproduces nn.BatchNorm2d on the global batch
     (= torch.nn.SyncBatchNorm semantics: var+eps, unbiased running_var over the global count);
  3. IoU — the integer intersection / union counts are all-reduced (not per-rank ratios), metrics.py:15-18 on the global
     batch; the scalar loss is averaged for logging.
With equal per-rank batches the mean of the local losses is the global-batch loss for both BCE (mean) and Dice
(per-sample mean), so averaging the local-loss gradients is exact."""
import torch
import torch.distributed as dist
from torch import nn


class GradBuckets(object):
    """Flat gradient buffer with bucketed asynchronous all-reduce (device-agnostic: tested with gloo on CPU)."""

    def __init__(self, params_in_backward_order, bucket_bytes=8 << 20, group=None):
        self.group = group
        self.params = list(params_in_backward_order)
        self.offsets, self.bucket_of, self.bucket_ranges = {}, {}, []
        off = start = 0
        for p in self.params:
            n = (p.numel() + 31) // 32 * 32          # keep every gradient 128-byte aligned
            self.offsets[id(p)] = (off, p.numel(), tuple(p.shape))
            self.bucket_of[id(p)] = len(self.bucket_ranges)
            off += n
            if (off - start) * 4 >= bucket_bytes:
                self.bucket_ranges.append((start, off))
                start = off
        if off > start:
            self.bucket_ranges.append((start, off))
        self.total = off
        self.flat = None
        self.pending = None
        self.works = []
        self.side = None  # CUDA stream(s) other than the caller's that gradients are produced on (engine: level / wgrad streams)

    def _launch(self, s, e):
        if self.side is not None and self.flat.is_cuda:
            cur = torch.cuda.current_stream(self.flat.device)   # the bucket holds gradients produced on the engine's other streams
            for st_ in (self.side if isinstance(self.side, (list, tuple)) else [self.side]):
                cur.wait_stream(st_)
        op = dist.ReduceOp.AVG if dist.get_backend(self.group) == "nccl" else dist.ReduceOp.SUM
        self.works.append((dist.all_reduce(self.flat[s:e], op=op, group=self.group, async_op=True), s, e, op))

    def begin(self, device):
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.pending = [0] * len(self.bucket_ranges)
        for p in self.params:
            self.pending[self.bucket_of[id(p)]] += 1
        self.works = []

    def view(self, p):
        off, n, shape = self.offsets[id(p)]
        return self.flat[off:off + n].view(shape)

    def ready(self, p):
        """Marks p's gradient as written; launches the bucket's all-reduce when it is complete."""
        b = self.bucket_of[id(p)]
        self.pending[b] -= 1
        if self.pending[b] == 0:
            self._launch(*self.bucket_ranges[b])

    def finish(self):
        """Flushes buckets whose parameters produced no gradient (kept zero) and waits (stream-level) for all transfers."""
        for b, left in enumerate(self.pending):
            if left > 0:
                self.pending[b] = 0
                self._launch(*self.bucket_ranges[b])
        ws = dist.get_world_size(self.group)
        for w, s, e, op in self.works:
            w.wait()
            if op == dist.ReduceOp.SUM:
                self.flat[s:e].div_(ws)
        self.works = []


class PeerExchange(object):
    """Peer-mapped exchange slots for the in-kernel SyncBN reduction (csrc/bn_tail.cuh): one symmetric allocation per rank
    (torch.distributed._symmetric_memory: CUDA VMM handles exchanged over the process group, every rank maps every peer's buffer),
    carved into per-(layer, direction) slots of [2 epoch parities][2][Cmax] fp64 sums plus 8 flag words, and a local epoch counter
    per slot.  Nothing here launches a collective: the conv / router kernels' last CTAs talk to each other through these pointers."""
    CMAX = 512

    def __init__(self, num_slots, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("the in-kernel SyncBN exchange covers one node (<= 8 ranks)")
        self.slot_bytes = 2 * 2 * self.CMAX * 8
        self.data_bytes = num_slots * self.slot_bytes
        total = self.data_bytes + num_slots * 8 * 4
        self.buf = symm.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.seq = torch.zeros(num_slots, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)          # every rank's slots and flags are zero before the first kernel touches a peer

    def slot(self, i):
        import _lib as L
        x = L.BnExchange()
        x.world, x.rank = self.world, self.rank
        for q in range(self.world):
            x.data[q] = self.ptrs[q] + i * self.slot_bytes
            x.flags[q] = self.ptrs[q] + self.data_bytes + i * 32
        x.seq = self.seq.data_ptr() + 4 * i
        return x


class DataParallel(nn.Module):
    """Wraps archs.UNet / archs.NestedUNet for one-process-per-GPU data parallelism.

    state_dict()/load_state_dict() keep the reference's key layout (no 'module.' prefix) so checkpoints interchange."""

    def __init__(self, module, sync_bn=True, bucket_bytes=8 << 20, process_group=None):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (backend nccl) before wrapping the model")
        self.module = module
        self.sync_bn = bool(sync_bn)
        self.group = process_group
        self.world_size = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=0, group=process_group)
        import engine
        order = []
        for s in reversed(engine._state(module).specs):
            if s.head is not None:
                order += [s.head.weight, s.head.bias]
            order += [s.bn.weight, s.bn.bias, s.conv.weight] + ([s.conv.bias] if s.conv.bias is not None else [])
        self.buckets = GradBuckets(order, bucket_bytes, process_group)
        # SyncBN inside the kernels that produce the sums (no collective launch, CUDA-graph capturable); NCCL all-reduce otherwise
        self.exchange = None
        import os
        dev = next(module.parameters()).device
        if self.sync_bn and self.world_size > 1 and dev.type == "cuda" and dist.get_backend(process_group) == "nccl" \
                and os.environ.get("NUNET_PEER_SYNCBN", "1") == "1":
            try:
                self.exchange = PeerExchange(2 * len(engine._state(module).specs), dev, process_group)
            except Exception as exc:    # no peer access / symmetric memory on this system: the NCCL path stays
                import warnings
                warnings.warn("nunet: peer-memory SyncBN exchange unavailable (%s); using NCCL all-reduces" % (exc,))
        module.__dict__["_nunet_dp"] = self

    # ---- engine hooks
    def all_reduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    # ---- nn.Module surface
    def forward(self, x):
        return self.module(x)

    def state_dict(self, *a, **kw):
        return self.module.state_dict(*a, **kw)

    def load_state_dict(self, *a, **kw):
        return self.module.load_state_dict(*a, **kw)


def global_iou_score(output, target, group=None):
    """metrics.iou_score over the GLOBAL batch: integer counts are all-reduced, then the float64 ratio is formed."""
    import metrics
    i, u = metrics.iou_counts(output, target)
    dev = output.device if torch.is_tensor(output) else "cpu"
    c = torch.tensor([i, u], dtype=torch.int64, device=dev)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    i, u = c.tolist()
    return (i + 1e-5) / (u + 1e-5)


def global_mean(value, device, group=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t /= dist.get_world_size(group)
    return float(t)
