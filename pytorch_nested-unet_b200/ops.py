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
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


# ---------------------------------------------------------------------------------------------- launch accounting
_LAUNCHES = 0
_TIMING = None  # list of (class, meta, start_event, end_event) while enabled


def reset_launch_count():
    global _LAUNCHES
    _LAUNCHES = 0


def launch_count():
    """Number of CUDA kernels of this library launched since reset_launch_count()."""
    return _LAUNCHES


def add_launches(n):
    """CUDA-graph replays launch the kernels that were counted while the graph was captured."""
    global _LAUNCHES
    _LAUNCHES += n


def kernel_timing_enabled():
    return _TIMING is not None


def enable_kernel_timing(on):
    """Per-launch CUDA-event timing on the launching (current) stream; used by bench.py for the roofline figures."""
    global _TIMING
    _TIMING = [] if on else None


def collect_kernel_timing():
    torch.cuda.synchronize()
    by_class, launches = {}, []
    for cls, meta, e0, e1 in (_TIMING or []):
        ms = e0.elapsed_time(e1)
        by_class[cls] = by_class.get(cls, 0.0) + ms
        launches.append({"class": cls, "meta": meta, "ms": ms})
    return {"by_class": by_class, "launches": launches}


def _call(cls, nkernels, meta, what, fn, *args):
    global _LAUNCHES
    if _TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        _TIMING.append((cls, meta, e0, e1))
    else:
        rc = fn(*args)
    L.check(rc, what)
    _LAUNCHES += nkernels


def require_cuda(t, what):
    if not t.is_cuda:
        raise L.NunetError("%s must be a CUDA tensor: the B200 path has no CPU fallback" % what)


def is_device_tensor(t):
    return t.is_cuda


@functools.lru_cache(maxsize=None)
def c32_eligible(N, H, W, cin_total, Cout):
    """True if a conv of this shape over pre-activated sources runs on the stacked Cout = 32 kernel (needs *_C32 weight packs)."""
    return bool(L.lib().nunet_conv3x3_c32_eligible(N, H, W, cin_total, Cout))


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
    _call("pack", 1, None, "nunet_pack_weights", L.lib().nunet_pack_weights, _ptr(w), w.shape[0], w.shape[1], mode, C.byref(plan), _ptr(out), _stream())
    return out


def pack_weights_batched(items):
    """items: [(fp32 OIHW weight, mode, plan, out bf16)] -> all packed by one launch per 64 layers."""
    jobs = (L.PackJob * len(items))()
    for j, (w, mode, plan, out) in zip(jobs, items):
        require_cuda(w, "weight")
        if w.dtype != torch.float32 or not w.is_contiguous():
            raise L.NunetError("pack_weights_batched expects contiguous fp32 OIHW weights")
        j.w, j.out, j.Cout, j.Cin, j.mode = w.data_ptr(), out.data_ptr(), w.shape[0], w.shape[1], mode
        j.NT, j.CK, j.nchunks, j.n_tiles = plan.NT, plan.CK, plan.nchunks, plan.n_tiles
    _call("pack", (len(items) + 63) // 64, None, "nunet_pack_weights_batched", L.lib().nunet_pack_weights_batched, C.cast(jobs, C.c_void_p), len(items), _stream())


def _p(t):
    return t.data_ptr() if t is not None else None


def bn_tail(counter, count, gamma, beta, conv_bias, running_mean, running_var, momentum, eps, scale, shift, save_mean, save_invstd, sums_out=None,
            xch=None):
    """nunet_bn_tail_t: BatchNorm finalisation in the tail of the conv that produced the statistics.  gamma None -> sums_out only.
    xch (L.BnExchange): cross-rank SyncBN inside the tail; `count` is then the global count."""
    t = L.BnTail()
    if xch is not None:
        t.xch = xch
    t.counter, t.count = counter.data_ptr(), float(count)
    t.gamma, t.beta, t.conv_bias = _p(gamma), _p(beta), _p(conv_bias)
    t.running_mean, t.running_var, t.momentum, t.eps = _p(running_mean), _p(running_var), float(momentum), float(eps)
    t.scale, t.shift, t.save_mean, t.save_invstd, t.sums_out = _p(scale), _p(shift), _p(save_mean), _p(save_invstd), _p(sums_out)
    return t


def bn_bwd_tail(counter, count, gamma, save_mean, save_invstd, dgamma, dbeta, k1, k2, k3, grad_scale=1.0, sums_out=None, xch=None):
    """nunet_bn_bwd_tail_t: BatchNorm-backward coefficients in the tail of the kernel that produced the two sums."""
    t = L.BnBwdTail()
    if xch is not None:
        t.xch = xch
    t.counter, t.count = counter.data_ptr(), float(count)
    t.gamma, t.save_mean, t.save_invstd = _p(gamma), _p(save_mean), _p(save_invstd)
    t.dgamma, t.dbeta, t.k1, t.k2, t.k3 = _p(dgamma), _p(dbeta), _p(k1), _p(k2), _p(k3)
    t.grad_scale, t.sums_out = float(grad_scale), _p(sums_out)
    return t


class Mask(object):
    """Fused block-internal gradient router of a dgrad launch: the producer's raw output and BatchNorm coefficients."""
    __slots__ = ("y", "scale", "shift", "mean", "invstd")

    def __init__(self, y, scale, shift, mean, invstd):
        self.y, self.scale, self.shift, self.mean, self.invstd = y, scale, shift, mean, invstd


def conv3x3(N, H, W, sources, Cout, wpack, out=None, stat_part=None, want_stats=False, cls=None, stacked=False, tail=None,
            out_segs=None, mask=None, bwd_tail=None):
    """Launches the fused implicit-GEMM conv. Returns (out bf16 [N,H,W,Cout] — or the list of per-segment tensors when out_segs
    (channel counts) is given —, stat_part [grid_m,2,Cout] or None)."""
    cin = sum(s.cpad for s in sources)
    plan = conv_plan(N, H, W, cin, Cout)
    dev = sources[0].data.device
    if out_segs is not None:
        out = [torch.empty((N, H, W, c), dtype=torch.bfloat16, device=dev) for c in out_segs]
    elif out is None:
        out = torch.empty((N, H, W, Cout), dtype=torch.bfloat16, device=dev)
    if mask is not None:
        want_stats = True
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
    d.wpack_layout = 1 if stacked else 0
    if out_segs is not None:
        d.num_out_seg = len(out)
        for i, (t, c) in enumerate(zip(out, out_segs)):
            d.out_seg[i], d.out_seg_c[i] = t.data_ptr(), c
    else:
        d.out = out.data_ptr()
    d.stat_part = stat_part.data_ptr() if stat_part is not None else None
    if tail is not None:
        d.bn = tail
    if mask is not None:
        d.mask_y, d.mask_scale, d.mask_shift = mask.y.data_ptr(), mask.scale.data_ptr(), mask.shift.data_ptr()
        d.mask_mean, d.mask_invstd = mask.mean.data_ptr(), mask.invstd.data_ptr()
        if bwd_tail is not None:
            d.bn_bwd = bwd_tail
    if cls is None:
        cls = "conv_dgrad" if sources[0].kind == L.SRC_GRAD else "conv_fwd"
    _call(cls, 1, (N, H, W, cin, Cout, "c32" if stacked else "igemm"), "nunet_conv3x3_fwd", L.lib().nunet_conv3x3_fwd, C.byref(d), _stream())
    return out, stat_part


class BnWorkspace(object):
    """Scratch shared by every BN finalize launch on one device (launches are stream-ordered)."""

    def __init__(self, device, cmax=512):
        self.scratch = torch.empty(L.lib().nunet_bn_workspace_bytes(cmax) // 8, dtype=torch.float64, device=device)
        self.counters = torch.zeros(cmax // 32, dtype=torch.int32, device=device)


def bn_finalize_train(stat_part, count, gamma, beta, conv_bias, running_mean, running_var, momentum, eps, ws,
                      scale, shift, save_mean, save_invstd, sums_out=None, finalize=True):
    rows, _, Cc = stat_part.shape
    _call("bn_finalize", 1, None, "nunet_bn_finalize_train", L.lib().nunet_bn_finalize_train, 
        _ptr(stat_part), rows, Cc, float(count), _ptr(gamma if finalize else None), _ptr(beta), _ptr(conv_bias),
        _ptr(running_mean), _ptr(running_var), float(momentum), float(eps), _ptr(scale), _ptr(shift), _ptr(save_mean),
        _ptr(save_invstd), _ptr(ws.scratch), _ptr(ws.counters), _ptr(sums_out), _stream())


def bn_finalize_from_sums(sums, count, gamma, beta, conv_bias, running_mean, running_var, momentum, eps, scale, shift,
                          save_mean, save_invstd):
    _call("bn_finalize", 1, None, "nunet_bn_finalize_from_sums", L.lib().nunet_bn_finalize_from_sums, 
        _ptr(sums), gamma.numel(), float(count), _ptr(gamma), _ptr(beta), _ptr(conv_bias), _ptr(running_mean),
        _ptr(running_var), float(momentum), float(eps), _ptr(scale), _ptr(shift), _ptr(save_mean), _ptr(save_invstd),
        _stream())


def bn_apply(y, scale, shift, out=None):
    """a = relu(scale*y+shift), bf16 NHWC -> bf16 NHWC (materialised once for node outputs that many loaders re-read)."""
    if out is None:
        out = torch.empty_like(y)
    Cc = y.shape[-1]
    _call("bn_apply", 1, None, "nunet_bn_apply", L.lib().nunet_bn_apply, _ptr(y), _ptr(scale), _ptr(shift), y.numel() // Cc, Cc, _ptr(out), _stream())
    return out


def bn_activate(y, scale, shift, want_a=True, want_pool=False, want_up=False, head=None):
    """One pass over a raw conv output y (bf16 NHWC): returns (a, a_pool, a_up, logits), each None unless requested.
    head = (wh [K,32] fp32, bh [K] fp32 or None): the 1x1 head's logits (fp32 NCHW) from the fp32 activation."""
    N, H, W, Cc = y.shape
    dev = y.device
    bf = dict(dtype=torch.bfloat16, device=dev)
    a = torch.empty((N, H, W, Cc), **bf) if want_a else None
    ap = torch.empty((N, H // 2, W // 2, Cc), **bf) if want_pool else None
    au = torch.empty((N, 2 * H, 2 * W, Cc), **bf) if want_up else None
    d = L.ActivateDesc()
    d.y, d.N, d.H, d.W, d.C, d.scale, d.shift = y.data_ptr(), N, H, W, Cc, scale.data_ptr(), shift.data_ptr()
    d.a, d.a_pool, d.a_up = _p(a), _p(ap), _p(au)
    logits = None
    if head is not None:
        wh, bh = head
        logits = torch.empty((N, wh.shape[0], H, W), dtype=torch.float32, device=dev)
        d.wh, d.bh, d.K, d.logits = wh.data_ptr(), _p(bh), wh.shape[0], logits.data_ptr()
    _call("activate", 1, (N, H, W, Cc, int(want_a), int(want_pool), int(want_up), int(head is not None)), "nunet_bn_activate",
          L.lib().nunet_bn_activate, C.byref(d), _stream())
    return a, ap, au, logits


def upsample2x(a, out=None):
    """bf16 NHWC [N,h,w,C] -> [N,2h,2w,C], bilinear align_corners=True."""
    N, h, w, Cc = a.shape
    if out is None:
        out = torch.empty((N, 2 * h, 2 * w, Cc), dtype=torch.bfloat16, device=a.device)
    _call("upsample2x", 1, None, "nunet_upsample2x", L.lib().nunet_upsample2x, _ptr(a), N, h, w, Cc, _ptr(out), _stream())
    return out


def maxpool2x2(a, out=None):
    """bf16 NHWC [N,H,W,C] -> [N,H/2,W/2,C] (nn.MaxPool2d(2, 2) of an activated node output)."""
    N, H, W, Cc = a.shape
    if out is None:
        out = torch.empty((N, H // 2, W // 2, Cc), dtype=torch.bfloat16, device=a.device)
    _call("maxpool2x2", 1, None, "nunet_maxpool2x2", L.lib().nunet_maxpool2x2, _ptr(a), N, H, W, Cc, _ptr(out), _stream())
    return out


def input_to_nhwc32(x, out=None):
    """fp32 NCHW network input -> bf16 NHWC with 32 channels (zero-padded)."""
    N, Cc, H, W = x.shape
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise L.NunetError("input_to_nhwc32 expects a contiguous fp32 NCHW tensor")
    if out is None:
        out = torch.empty((N, H, W, 32), dtype=torch.bfloat16, device=x.device)
    _call("input_pack", 1, None, "nunet_input_to_nhwc32", L.lib().nunet_input_to_nhwc32, _ptr(x), N, Cc, H, W, _ptr(out), _stream())
    return out


IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)   # albumentations Normalize() defaults (trains.py:266)
_NORM_CACHE = {}


def _norm_tensors(mean, std, C_, dev):
    key = (tuple(mean), tuple(std), C_, dev)
    t = _NORM_CACHE.get(key)
    if t is None:
        m = torch.tensor([mean[i % len(mean)] for i in range(C_)], dtype=torch.float32, device=dev)
        s_ = torch.tensor([std[i % len(std)] for i in range(C_)], dtype=torch.float32, device=dev)
        t = _NORM_CACHE[key] = (m, s_)
    return t


def input_u8(img, mean=IMAGENET_MEAN, std=IMAGENET_STD, max_pixel=255.0, post_div=255.0, aug=None, nhwc32=True):
    """uint8 NHWC image batch -> the network input: bf16 NHWC32 (stem operand) or fp32 NCHW; aug: optional uint8 [N] (see header)."""
    require_cuda(img, "image batch")
    if img.dtype != torch.uint8 or img.dim() != 4 or not img.is_contiguous():
        raise L.NunetError("input_u8 expects a contiguous uint8 NHWC tensor")
    N, H, W, Cc = img.shape
    m, s_ = _norm_tensors(mean, std, Cc, img.device)
    out = torch.empty((N, H, W, 32), dtype=torch.bfloat16, device=img.device) if nhwc32 else torch.empty((N, Cc, H, W), dtype=torch.float32, device=img.device)
    _call("input_pack", 1, None, "nunet_input_u8", L.lib().nunet_input_u8, _ptr(img), N, H, W, Cc, _ptr(m), _ptr(s_), float(max_pixel), float(post_div),
          _ptr(aug), _ptr(out), 0 if nhwc32 else 1, _stream())
    return out


def mask_u8(mask, max_pixel=255.0, aug=None):
    """uint8 NHWC mask stack [N,H,W,K] -> fp32 NCHW mask/255 (dataset.py:73-74), with the same geometric augmentation as the image."""
    require_cuda(mask, "mask batch")
    if mask.dtype != torch.uint8 or mask.dim() != 4 or not mask.is_contiguous():
        raise L.NunetError("mask_u8 expects a contiguous uint8 NHWC tensor")
    N, H, W, K = mask.shape
    out = torch.empty((N, K, H, W), dtype=torch.float32, device=mask.device)
    _call("input_pack", 1, None, "nunet_input_u8", L.lib().nunet_input_u8, _ptr(mask), N, H, W, K, _ptr(None), _ptr(None), float(max_pixel), 1.0,
          _ptr(aug), _ptr(out), 1, _stream())
    return out


def bn_fold_eval(gamma, beta, conv_bias, running_mean, running_var, eps, scale, shift):
    _call("bn_fold", 1, None, "nunet_bn_fold_eval", L.lib().nunet_bn_fold_eval, gamma.numel(), _ptr(gamma), _ptr(beta), _ptr(conv_bias), _ptr(running_mean),
                                       _ptr(running_var), float(eps), _ptr(scale), _ptr(shift), _stream())


# ------------------------------------------------------------------------------------------------ heads / loss / metric
def head_fwd(y, scale, shift, wh, bh, N, H, W, logits=None):
    K = wh.shape[0]
    if logits is None:
        logits = torch.empty((N, K, H, W), dtype=torch.float32, device=y.device)
    _call("head_fwd", 1, None, "nunet_head_fwd", L.lib().nunet_head_fwd, _ptr(y), _ptr(scale), _ptr(shift), _ptr(wh), _ptr(bh), N, H, W, K, _ptr(logits), _stream())
    return logits


def loss_iou_fwd(logits, target, want_counts=True):
    """Returns (loss 0-dim fp32 tensor, sums [N,3] fp64, counts [2] int64 or None). No host sync."""
    N = logits.shape[0]
    E = logits.numel() // N
    bps = L.lib().nunet_loss_blocks_per_sample(N, E)
    dev = logits.device
    part = torch.empty((N, bps, 4), dtype=torch.float32, device=dev)
    cnt = torch.empty((N, bps, 2), dtype=torch.int32, device=dev)
    sums = torch.empty((N, 3), dtype=torch.float64, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    counts = torch.empty(2, dtype=torch.int64, device=dev) if want_counts else None
    _call("loss_fwd", 2, None, "nunet_loss_iou_fwd", L.lib().nunet_loss_iou_fwd, _ptr(logits), _ptr(target), N, E, _ptr(part), _ptr(cnt), _ptr(sums), _ptr(loss), _ptr(counts),
                                       _stream())
    return loss, sums, counts


def iou_counts(logits, target):
    counts = torch.empty(2, dtype=torch.int64, device=logits.device)
    _call("iou", 1, None, "nunet_iou_counts", L.lib().nunet_iou_counts, _ptr(logits), _ptr(target), logits.numel(), _ptr(counts), _stream())
    return counts


def lovasz_hinge_fwd(logits, target):
    """Returns (loss 0-dim fp32, coef [N,P] fp32 = d loss_image / d logit).  logits / target: [N, ...] fp32, flattened per image."""
    N = logits.shape[0]
    P = logits.numel() // N
    dev = logits.device
    n_scr = L.lib().nunet_lovasz_scratch_elems(N, P)
    scratch = torch.empty(n_scr, dtype=torch.int64, device=dev) if n_scr else None
    coef = torch.empty((N, P), dtype=torch.float32, device=dev)
    limg = torch.empty(N, dtype=torch.float64, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    _call("loss_fwd", 2, None, "nunet_lovasz_hinge_fwd", L.lib().nunet_lovasz_hinge_fwd, _ptr(logits), _ptr(target), N, P, _ptr(scratch), _ptr(coef),
          _ptr(limg), _ptr(loss), _stream())
    return loss, coef


def lovasz_hinge_bwd(coef, gscale, gmul, shape):
    d = torch.empty(shape, dtype=torch.float32, device=coef.device)
    _call("loss_bwd", 1, None, "nunet_lovasz_hinge_bwd", L.lib().nunet_lovasz_hinge_bwd, _ptr(coef), _ptr(gscale), float(gmul), coef.numel(), _ptr(d), _stream())
    return d


def sigmoid_u8(logits):
    """(sigmoid(logits) * 255).astype('uint8') on the device (val.py:100-105), same shape."""
    require_cuda(logits, "logits")
    x = logits.detach()
    x = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous()
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    _call("val_out", 1, None, "nunet_sigmoid_u8", L.lib().nunet_sigmoid_u8, _ptr(x), x.numel(), _ptr(out), _stream())
    return out


def loss_bwd(logits, target, sums, gscale, gmul=1.0, dlogits=None):
    N = logits.shape[0]
    E = logits.numel() // N
    if dlogits is None:
        dlogits = torch.empty_like(logits)
    _call("loss_bwd", 1, None, "nunet_loss_bwd", L.lib().nunet_loss_bwd, _ptr(logits), _ptr(target), _ptr(sums), _ptr(gscale), float(gmul), N, E, _ptr(dlogits), _stream())
    return dlogits


MAX_LOSS_HEADS = 4


def _ptr_array(tensors):
    """Host array of device pointers (None -> NULL) for the *_multi entry points."""
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def loss_iou_fwd_multi(logits_list, target):
    """BCEDice + IoU counts of nheads <= 4 logit tensors against ONE target in one pass (two launches instead of 2 * nheads).
    Returns (loss [nheads] fp32, sums [nheads,N,3] fp64, counts [nheads,2] int64); bitwise loss_iou_fwd per head."""
    nh = len(logits_list)
    N = logits_list[0].shape[0]
    E = logits_list[0].numel() // N
    bps = L.lib().nunet_loss_blocks_per_sample(N, E)
    dev = logits_list[0].device
    part = torch.empty((nh, N, bps, 4), dtype=torch.float32, device=dev)
    cnt = torch.empty((nh, N, bps, 2), dtype=torch.int32, device=dev)
    sums = torch.empty((nh, N, 3), dtype=torch.float64, device=dev)
    loss = torch.empty(nh, dtype=torch.float32, device=dev)
    counts = torch.empty((nh, 2), dtype=torch.int64, device=dev)
    arr = _ptr_array(logits_list)
    _call("loss_fwd", 2, None, "nunet_loss_iou_fwd_multi", L.lib().nunet_loss_iou_fwd_multi, arr, nh, _ptr(target), N, E, _ptr(part), _ptr(cnt), _ptr(sums),
          _ptr(loss), _ptr(counts), _stream())
    return loss, sums, counts


def loss_bwd_multi(logits_list, target, sums, gscales, gmul=1.0):
    """dlogits of every head whose loss received a gradient (gscales[k] is its 0-dim upstream gradient, None = no gradient) in one
    launch; returns a list with None for the heads without gradient."""
    nh = len(logits_list)
    N = logits_list[0].shape[0]
    E = logits_list[0].numel() // N
    outs = [torch.empty_like(l) if g is not None else None for l, g in zip(logits_list, gscales)]
    if all(o is None for o in outs):
        return outs
    la, ga, da = _ptr_array(logits_list), _ptr_array(gscales), _ptr_array(outs)
    _call("loss_bwd", 1, None, "nunet_loss_bwd_multi", L.lib().nunet_loss_bwd_multi, la, nh, _ptr(target), _ptr(sums), ga, float(gmul), N, E, da, _stream())
    return outs


def head_bwd_params(y, scale, shift, dlogits, N, H, W, K, dwh=None, dbh=None):
    dev = y.device
    rows = L.lib().nunet_head_bwd_params_rows()
    part = torch.empty((rows, K, 33), dtype=torch.float32, device=dev)
    if dwh is None:
        dwh = torch.empty((K, 32), dtype=torch.float32, device=dev)
    if dbh is None:
        dbh = torch.empty((K,), dtype=torch.float32, device=dev)
    _call("head_bwd", 2, None, "nunet_head_bwd_params", L.lib().nunet_head_bwd_params, _ptr(y), _ptr(scale), _ptr(shift), _ptr(dlogits), N, H, W, K, _ptr(part), _ptr(dwh), _ptr(dbh), 0,
                                          _stream())
    return dwh, dbh


# ------------------------------------------------------------------------------------------------ backward
class Consumer(object):
    __slots__ = ("dx", "pitch", "coff", "kind")

    def __init__(self, dx, pitch, coff, kind):
        self.dx, self.pitch, self.coff, self.kind = dx, pitch, coff, kind


def grad_router(N, H, W, Cc, y, scale, shift, mean, invstd, consumers, gz=None, part=None, dlogits=None, wh=None, tail=None, head_grads=None):
    """head_grads = (dwh [K,32], dbh [K] or None): also produce the head's parameter gradients in this pass (K == 1 only; needs tail)."""
    dev = y.device
    rows = L.lib().nunet_grad_router_rows(N, H, W, Cc)
    if gz is None:
        gz = torch.empty((N, H, W, Cc), dtype=torch.bfloat16, device=dev)
    if part is None:
        part = torch.empty((rows, 2, Cc), dtype=torch.float32, device=dev)
    d = L.RouterDesc()
    d.N, d.H, d.W, d.C = N, H, W, Cc
    d.y, d.scale, d.shift, d.mean, d.invstd = y.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr()
    d.num_cons = len(consumers)
    for i, c in enumerate(consumers):
        d.cons[i].dx, d.cons[i].pitch, d.cons[i].coff, d.cons[i].kind = c.dx.data_ptr(), c.pitch, c.coff, c.kind
    if dlogits is not None:
        d.dlogits, d.wh, d.K = dlogits.data_ptr(), wh.data_ptr(), wh.shape[0]
    d.gz, d.part = gz.data_ptr(), part.data_ptr()
    if tail is not None:
        d.bn = tail
    if head_grads is not None:
        hp = torch.empty((rows, 33), dtype=torch.float32, device=dev)
        d.head_part, d.dwh, d.dbh = hp.data_ptr(), head_grads[0].data_ptr(), _p(head_grads[1])
    _call("router", 1, (N, H, W, Cc, len(consumers)), "nunet_grad_router", L.lib().nunet_grad_router, C.byref(d), _stream())
    return gz, part


def up_transpose(dx, pitch, coff, N, h, w, Cc, out=None):
    """Backward of the bilinear-2x loader: dx [N,2h,2w,pitch] slice -> dense bf16 [N,h,w,Cc]."""
    if out is None:
        out = torch.empty((N, h, w, Cc), dtype=torch.bfloat16, device=dx.device)
    _call("up_transpose", 1, (N, h, w, Cc), "nunet_up_transpose", L.lib().nunet_up_transpose, _ptr(dx), pitch, coff, N, h, w, Cc, _ptr(out), _stream())
    return out


def bn_backward_reduce(part, count, gamma, mean, invstd, ws, dgamma, dbeta, k1, k2, k3, sums_out=None, finalize=True):
    rows, _, Cc = part.shape
    _call("bn_bwd", 1, None, "nunet_bn_backward_reduce", L.lib().nunet_bn_backward_reduce, _ptr(part), rows, Cc, float(count), _ptr(gamma if finalize else None), _ptr(mean), _ptr(invstd),
                                             _ptr(dgamma), _ptr(dbeta), _ptr(k1), _ptr(k2), _ptr(k3), _ptr(ws.scratch), _ptr(ws.counters),
                                             _ptr(sums_out), _stream())


def bn_backward_from_sums(sums, count, gamma, mean, invstd, dgamma, dbeta, k1, k2, k3, grad_scale=1.0):
    _call("bn_bwd", 1, None, "nunet_bn_backward_from_sums", L.lib().nunet_bn_backward_from_sums, _ptr(sums), gamma.numel(), float(count), _ptr(gamma), _ptr(mean), _ptr(invstd), _ptr(dgamma),
                                                _ptr(dbeta), _ptr(k1), _ptr(k2), _ptr(k3), float(grad_scale), _stream())


@functools.lru_cache(maxsize=None)
def wgrad_plan(N, H, W, cin_total, Cout):
    pl = L.WgradPlan()
    L.check(L.lib().nunet_conv3x3_wgrad_plan(N, H, W, cin_total, Cout, C.byref(pl)), "nunet_conv3x3_wgrad_plan")
    return pl


def bn_bwd_apply(gz, y, k1, k2, k3, out=None):
    """g_y = k1*g_z + k2*y + k3 (bf16 NHWC); out may be gz itself (in place)."""
    if out is None:
        out = torch.empty_like(gz)
    Cc = gz.shape[-1]
    _call("bn_bwd_apply", 1, None, "nunet_bn_bwd_apply", L.lib().nunet_bn_bwd_apply, _ptr(gz), _ptr(y), _ptr(k1), _ptr(k2), _ptr(k3),
          gz.numel() // Cc, Cc, _ptr(out), _stream())
    return out


def conv3x3_wgrad(N, H, W, sources, Cout, gy, dw, part=None, cin_real=0):
    """dw: fp32 [Cout, Cin_real, 3, 3] (overwritten); gy from bn_bwd_apply."""
    cin = sum(s.cpad for s in sources)
    plan = wgrad_plan(N, H, W, cin, Cout)
    if part is None:
        part = torch.empty(plan.part_elems, dtype=torch.float32, device=gy.device)
    d = L.WgradDesc()
    d.N, d.H, d.W, d.num_src, d.Cout = N, H, W, len(sources), Cout
    for i, s in enumerate(sources):
        d.src[i].data = s.data.data_ptr()
        d.src[i].scale = s.scale.data_ptr() if s.scale is not None else None
        d.src[i].shift = s.shift.data_ptr() if s.shift is not None else None
        d.src[i].C = s.C
        d.src[i].kind = s.kind
    d.gy = gy.data_ptr()
    d.part, d.dw = part.data_ptr(), dw.data_ptr()
    d.cin_real = cin_real
    _call("conv_wgrad", 2, (N, H, W, cin, Cout), "nunet_conv3x3_wgrad", L.lib().nunet_conv3x3_wgrad, C.byref(d), _stream())
    return dw
