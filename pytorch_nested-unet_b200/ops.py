"""Tensor-level wrappers over the C ABI: torch only provides device memory and the current stream.

Layouts: raw conv outputs are bf16 NHWC [N,H,W,C]; the network input stays fp32 NCHW (read directly by the stem conv's
loader); parameters are fp32 (OIHW) masters that `pack_weights` turns into the bf16 tile images the kernels consume."""
import ctypes as C
import functools

import torch

import _lib as L


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, what):
    if not t.is_cuda:
        raise L.NunetError("%s must be a CUDA tensor: the B200 path has no CPU fallback" % what)


@functools.lru_cache(maxsize=None)
def conv_plan(N, H, W, cin_total, Cout):
    pl = L.ConvPlan()
    L.check(L.lib().nunet_conv3x3_plan(N, H, W, cin_total, Cout, C.byref(pl)), "nunet_conv3x3_plan(%s)" % ((N, H, W, cin_total, Cout),))
    return pl


class Source(object):
    """One concatenated source of a fused conv (kind = L.SRC_*)."""
    __slots__ = ("kind", "data", "data2", "scale", "shift", "k2", "C")

    def __init__(self, kind, data, C_, scale=None, shift=None, data2=None, k2=None):
        self.kind, self.data, self.C, self.scale, self.shift, self.data2, self.k2 = kind, data, C_, scale, shift, data2, k2

    @property
    def cpad(self):
        return 16 if self.kind == L.SRC_INPUT else self.C


def pack_weights(w_oihw, mode, plan, out=None):
    """fp32 OIHW -> packed bf16 (one launch)."""
    require_cuda(w_oihw, "weight")
    w = w_oihw.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    if out is None:
        out = torch.empty(plan.wpack_elems, dtype=torch.bfloat16, device=w.device)
    L.check(L.lib().nunet_pack_weights(_ptr(w), w.shape[0], w.shape[1], mode, C.byref(plan), _ptr(out), _stream()), "nunet_pack_weights")
    return out


def conv3x3(N, H, W, sources, Cout, wpack, out=None, stat_part=None, want_stats=False):
    """Launches the fused implicit-GEMM conv. Returns (out bf16 [N,H,W,Cout], stat_part [grid_m,2,Cout] or None)."""
    cin = sum(s.cpad for s in sources)
    plan = conv_plan(N, H, W, cin, Cout)
    dev = sources[0].data.device
    if out is None:
        out = torch.empty((N, H, W, Cout), dtype=torch.bfloat16, device=dev)
    if want_stats and stat_part is None:
        stat_part = torch.empty((plan.grid_m, 2, Cout), dtype=torch.float32, device=dev)
    d = L.ConvDesc()
    d.N, d.H, d.W, d.num_src, d.Cout = N, H, W, len(sources), Cout
    for i, s in enumerate(sources):
        d.src[i].data = s.data.data_ptr()
        d.src[i].data2 = s.data2.data_ptr() if s.data2 is not None else None
        d.src[i].scale = s.scale.data_ptr() if s.scale is not None else None
        d.src[i].shift = s.shift.data_ptr() if s.shift is not None else None
        d.src[i].k2 = s.k2.data_ptr() if s.k2 is not None else None
        d.src[i].C = s.C
        d.src[i].kind = s.kind
    d.wpack = wpack.data_ptr()
    d.out = out.data_ptr()
    d.stat_part = stat_part.data_ptr() if stat_part is not None else None
    L.check(L.lib().nunet_conv3x3_fwd(C.byref(d), _stream()), "nunet_conv3x3_fwd")
    return out, stat_part


class BnWorkspace(object):
    """Scratch shared by every BN finalize launch on one device (launches are stream-ordered)."""

    def __init__(self, device, cmax=512):
        self.scratch = torch.empty(L.lib().nunet_bn_workspace_bytes(cmax) // 8, dtype=torch.float64, device=device)
        self.counters = torch.zeros(cmax // 32, dtype=torch.int32, device=device)


def bn_finalize_train(stat_part, count, gamma, beta, conv_bias, running_mean, running_var, momentum, eps, ws,
                      scale, shift, save_mean, save_invstd, sums_out=None, finalize=True):
    rows, _, Cc = stat_part.shape
    L.check(L.lib().nunet_bn_finalize_train(
        _ptr(stat_part), rows, Cc, float(count), _ptr(gamma if finalize else None), _ptr(beta), _ptr(conv_bias),
        _ptr(running_mean), _ptr(running_var), float(momentum), float(eps), _ptr(scale), _ptr(shift), _ptr(save_mean),
        _ptr(save_invstd), _ptr(ws.scratch), _ptr(ws.counters), _ptr(sums_out), _stream()), "nunet_bn_finalize_train")


def bn_finalize_from_sums(sums, count, gamma, beta, conv_bias, running_mean, running_var, momentum, eps, scale, shift,
                          save_mean, save_invstd):
    L.check(L.lib().nunet_bn_finalize_from_sums(
        _ptr(sums), gamma.numel(), float(count), _ptr(gamma), _ptr(beta), _ptr(conv_bias), _ptr(running_mean),
        _ptr(running_var), float(momentum), float(eps), _ptr(scale), _ptr(shift), _ptr(save_mean), _ptr(save_invstd),
        _stream()), "nunet_bn_finalize_from_sums")


def bn_fold_eval(gamma, beta, conv_bias, running_mean, running_var, eps, scale, shift):
    L.check(L.lib().nunet_bn_fold_eval(gamma.numel(), _ptr(gamma), _ptr(beta), _ptr(conv_bias), _ptr(running_mean),
                                       _ptr(running_var), float(eps), _ptr(scale), _ptr(shift), _stream()), "nunet_bn_fold_eval")
