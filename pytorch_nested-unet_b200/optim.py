"""Drop-in replacements for the two optimizers the reference's training script constructs (trains.py:225-233): torch.optim.SGD
and torch.optim.Adam, with `step()` as ONE fused multi-tensor launch on the fp32 master parameters (csrc/optim.cu) instead of
torch's foreach path (~10 launches + host-side list handling over 128 tensors per step).

They subclass the torch optimizers: `param_groups`, `state_dict()` / `load_state_dict()`, `zero_grad()`, `add_param_group` and
every torch lr_scheduler (CosineAnnealingLR, ReduceLROnPlateau, MultiStepLR: trains.py:237-248) work unchanged, and the state
(`momentum_buffer`, `exp_avg`, `exp_avg_sq`, `step`) is stored under torch's keys.  Hyper-parameters are read by the kernel from a
small device tensor, so a captured CUDA graph follows the schedule.  The arithmetic is torch's, operation for operation
(tests/test_optim_gpu.py: bitwise for SGD, <= 2 ulp for Adam's division / square root).
"""
import ctypes as C

import numpy as np
import torch

import _lib as L
import ops


class _FusedMixin(object):
    _algo = None

    def _fused_init(self):
        self._plans = {}     # group index -> cached device tables

    def _group_tables(self, gi, group, states, first=None):
        """(tensor table, chunk table, nchunks) on the device for this group; rebuilt only when a pointer changed."""
        params = group["params"]
        key = tuple((p.data_ptr(), p.grad.data_ptr() if p.grad is not None else 0) + tuple(s.data_ptr() if s is not None else 0 for s in st)
                    for p, st in zip(params, states)) + (tuple(first) if first and any(first) else ())
        hit = self._plans.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1:4]
        dev = params[0].device
        chunk = L.lib().nunet_optim_chunk_elems()
        if hit is not None and hit[4] is not None:
            hit[4].synchronize()                 # the previous upload has left the pinned staging buffer
            stage, tab_dev, chunks_dev, nch = hit[5], hit[1], hit[2], hit[3]
        else:
            stage = torch.zeros((len(params), 5), dtype=torch.int64).pin_memory()
            tab_dev = chunks_dev = None
        tab = stage.numpy()
        chunks = []
        for i, (p, st) in enumerate(zip(params, states)):
            if p.dtype != torch.float32 or not p.is_contiguous() or (p.grad is not None and (p.grad.dtype != torch.float32 or not p.grad.is_contiguous())):
                raise L.NunetError("the fused optimizers expect contiguous fp32 parameters and gradients")
            tab[i] = (p.data_ptr(), p.grad.data_ptr() if p.grad is not None else 0, st[0].data_ptr() if st[0] is not None else 0,
                      st[1].data_ptr() if st[1] is not None else 0, p.numel() | ((1 << 62) if (first and first[i]) else 0))
            chunks += [(i, c) for c in range((p.numel() + chunk - 1) // chunk)]
        if tab_dev is None:
            tab_dev = torch.empty((len(params), 5), dtype=torch.int64, device=dev)
            chunks_dev = torch.tensor(chunks, dtype=torch.int32).to(dev)     # depends on the parameter sizes only
        tab_dev.copy_(stage, non_blocking=True)   # stream-ordered before the launch that reads it; no host synchronisation
        ev = torch.cuda.Event()
        ev.record()
        self._plans[gi] = (key, tab_dev, chunks_dev, len(chunks), ev, stage)
        return tab_dev, chunks_dev, len(chunks)

    def _hyper(self, gi, values, dev):
        """Device copy of the group's hyper-parameters, rewritten only when a value changed (a learning-rate scheduler stepped)."""
        slot = self.__dict__.setdefault("_hyper_dev", {})
        hit = slot.get(gi)
        vals = tuple(float(v) for v in values)
        if hit is None or hit[0].device != dev:
            hit = slot[gi] = [torch.tensor(vals, dtype=torch.float32, device=dev), vals]
        elif hit[1] != vals:
            hit[0].copy_(torch.tensor(vals, dtype=torch.float32), non_blocking=True)
            hit[1] = vals
        return hit[0]

    def _launch(self, tab, chunks, n, hyper, step_dev=None):
        ops._call("optimizer", 1 if step_dev is None else 2, None, "nunet_optim_step", L.lib().nunet_optim_step, C.c_void_p(tab.data_ptr()),
                  tab.shape[0], C.c_void_p(chunks.data_ptr()), n, C.c_void_p(hyper.data_ptr()), self._algo,
                  C.c_void_p(step_dev.data_ptr()) if step_dev is not None else C.c_void_p(0), ops._stream())
        import engine
        engine.weights_updated()     # the packed bf16 weight images are stale now


class SGD(_FusedMixin, torch.optim.SGD):
    _algo = 0

    def __init__(self, params, lr=1e-3, momentum=0, dampening=0, weight_decay=0, nesterov=False, **kw):
        torch.optim.SGD.__init__(self, params, lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov, **kw)
        self._fused_init()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            if group.get("maximize"):
                raise L.NunetError("maximize=True is not supported by the fused SGD")
            params = [p for p in group["params"]]
            if not params or not params[0].is_cuda:
                raise L.NunetError("the fused optimizers run on CUDA parameters only (no CPU fallback)")
            states, first = [], []
            for p in params:
                st = self.state[p]
                fresh = False
                if group["momentum"] != 0 and st.get("momentum_buffer") is None and p.grad is not None:
                    st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    fresh = group["dampening"] != 0      # torch: buf = clone(grad) on the first step (== 0*momentum + grad unless dampened)
                first.append(fresh)
                states.append((st.get("momentum_buffer") if group["momentum"] != 0 else None, None))
            tab, chunks, n = self._group_tables(gi, group, states, first)
            hyper = self._hyper(gi, (group["lr"], group["momentum"], group["dampening"], group["weight_decay"], 1.0 if group["nesterov"] else 0.0), params[0].device)
            self._launch(tab, chunks, n, hyper)
        return loss


class Adam(_FusedMixin, torch.optim.Adam):
    _algo = 1

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, **kw):
        torch.optim.Adam.__init__(self, params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, **kw)
        self._fused_init()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            if group.get("amsgrad") or group.get("maximize"):
                raise L.NunetError("amsgrad / maximize are not supported by the fused Adam")
            params = [p for p in group["params"]]
            if not params or not params[0].is_cuda:
                raise L.NunetError("the fused optimizers run on CUDA parameters only (no CPU fallback)")
            dev = params[0].device
            states = []
            for p in params:
                st = self.state[p]
                if "exp_avg" not in st:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                states.append((st["exp_avg"], st["exp_avg_sq"]))
            step_dev = self.__dict__.setdefault("_step_dev", {}).get(gi)
            if step_dev is None or step_dev.device != dev:     # one counter per parameter, seeded from torch's (host) state
                step_dev = self._step_dev[gi] = torch.tensor([int(self.state[p]["step"]) for p in params], dtype=torch.int32, device=dev)
            tab, chunks, n = self._group_tables(gi, group, states)
            b1, b2 = group["betas"]
            hyper = self._hyper(gi, (group["lr"], b1, b2, group["eps"], group["weight_decay"]), dev)
            self._launch(tab, chunks, n, hyper, step_dev)
            for p in params:                     # torch's per-parameter step counters (host tensors; kept for state_dict compatibility)
                if p.grad is not None:
                    self.state[p]["step"] += 1
        return loss
