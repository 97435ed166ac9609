"""Host-side schedule of the fused UNet++ / U-Net step on one B200.

`run_model(model, x)` walks the reference's node graph (finished/archs1.py:58-71 / :114-143) in its topological
order and issues, per VGGBlock conv, ONE fused implicit-GEMM launch (ops.conv3x3: the producers' BatchNorm+ReLU, the
2x2 max-pool / bilinear 2x upsample of the skip pathway and the channel concat all happen in its loader; train-mode
batch statistics come out of its epilogue) followed by one tiny BN finalize launch.  Activations exist in HBM only as
raw bf16 NHWC conv outputs.  The backward pass (an autograd.Function so that `loss.backward()` and torch.optim keep
working on the fp32 master parameters) runs, per conv in reverse order: gradient router + BN/ReLU backward reduction,
split-K wgrad, dgrad — all CUDA kernels behind the C ABI; nothing here calls an ATen compute op on activations.
"""
import os as _os
import weakref as _weakref

import torch

import _lib as L
import ops

NB_FILTER = (32, 64, 128, 256, 512)
# Node outputs (up to 5 forward + 5 wgrad loaders re-read each) are activated once into a bf16 tensor; False = every
# loader applies BatchNorm+ReLU itself from the raw conv output.
MATERIALIZE_NODE_OUTPUTS = True
# Also a block's first conv output (one forward + one wgrad reader): every conv then runs in copy mode and every weight
# gradient through the TMA-fed kernel (conv_wgrad_tma.cu), which more than pays for the extra write + read.
MATERIALIZE_ALL = _os.environ.get("NUNET_MATERIALIZE_ALL", "1") == "1"
# 2x2 max-pooled backbone outputs are written once (quarter size) instead of being pooled in the consumer's loaders.
MATERIALIZE_POOLED = _os.environ.get("NUNET_MATERIALIZE_POOLED", "1") == "1"
# The fp32 NCHW network input is repacked once per step to bf16 NHWC with 32 (zero-padded) channels: the stem conv becomes a
# plain 32-channel copy-mode layer in forward and weight gradient.
PACK_INPUT = _os.environ.get("NUNET_PACK_INPUT", "1") == "1"
# The upsampled skip operand is re-read by the forward conv and by Cout/32 wgrad CTAs per position; blending it in every
# loader costs 4-8x the arithmetic of writing it once (measured: tools/prof_wgrad.py). False = fused NUNET_SRC_UP* loaders.
MATERIALIZE_UPSAMPLED = True
WGRAD_SIDE_STREAM = _os.environ.get("NUNET_WGRAD_SIDE_STREAM", "1") == "1"
# The gradient router of a block's first conv (one consumer: the block's second conv) runs in the epilogue of that conv's dgrad.
FUSE_BLOCK_ROUTER = _os.environ.get("NUNET_FUSE_BLOCK_ROUTER", "1") == "1"
# Training steps run every resolution level of the nested graph on its own CUDA stream (+ one weight-gradient stream per
# level): nodes of one anti-diagonal are independent, so the HBM-bound kernels of one level (BN apply, upsample, gradient
# router) execute while another level's convolution holds the tensor pipes.  Cross-level edges (pool / upsample in forward,
# their dgrad slices in backward) are CUDA events; all streams fork from and join the caller's stream at the step boundaries,
# and every tensor of the step stays referenced until the join is enqueued, so no block is recycled under a reader.
LEVEL_STREAMS = _os.environ.get("NUNET_LEVEL_STREAMS", "1") == "1"


# --------------------------------------------------------------------------------------------- graphs
def nested_graph():
    """[(node, [(source node | 'input', kind)])] in the execution order of NestedUNet.forward (finished/archs1.py:115-132);
    concat order = same-level nodes in increasing j, the upsampled lower node last."""
    g = [("0_0", [("input", "input")])]
    for name in ("1_0", "0_1", "2_0", "1_1", "0_2", "3_0", "2_1", "1_2", "0_3", "4_0", "3_1", "2_2", "1_3", "0_4"):
        i, j = int(name[0]), int(name[2])
        if j == 0:
            g.append((name, [("%d_0" % (i - 1), "pool")]))
        else:
            g.append((name, [("%d_%d" % (i, jj), "same") for jj in range(j)] + [("%d_%d" % (i + 1, j - 1), "up")]))
    return g


def unet_graph():
    """UNet.forward (finished/archs1.py:59-68)."""
    g = [("0_0", [("input", "input")])]
    for i in range(1, 5):
        g.append(("%d_0" % i, [("%d_0" % (i - 1), "pool")]))
    for name, same, up in (("3_1", "3_0", "4_0"), ("2_2", "2_0", "3_1"), ("1_3", "1_0", "2_2"), ("0_4", "0_0", "1_3")):
        g.append((name, [(same, "same"), (up, "up")]))
    return g


_KIND = {"input": L.SRC_INPUT, "same": L.SRC_SAME, "pool": L.SRC_POOL, "up": L.SRC_UP}


class LayerSpec(object):
    """One 3x3 conv (+ its BatchNorm) of the network."""
    __slots__ = ("idx", "node", "which", "level", "conv", "bn", "cin_real", "cin_pad", "cout", "sources", "consumers", "head")

    def __init__(self):
        self.sources = []    # [(producer layer idx | -1, kind, C)]
        self.consumers = []  # [(consumer layer idx, channel offset in its concatenated input, kind, position in its source list)]
        self.head = None     # nn.Conv2d 1x1 reading this layer's activation


def build_specs(model):
    specs, node_out = [], {}
    for node, srcs in model._graph:
        block = getattr(model, "conv" + node)
        level = int(node[0])
        for which in (1, 2):
            s = LayerSpec()
            s.idx, s.node, s.which, s.level = len(specs), node, which, level
            s.conv = getattr(block, "conv%d" % which)
            s.bn = getattr(block, "bn%d" % which)
            s.cout = s.conv.out_channels
            s.cin_real = s.conv.in_channels
            if which == 1:
                for src, kind in srcs:
                    if kind == "input":
                        s.sources.append((-1, L.SRC_INPUT, s.cin_real))
                    else:
                        pi = node_out[src]
                        s.sources.append((pi, _KIND[kind], specs[pi].cout))
            else:
                s.sources.append((s.idx - 1, L.SRC_SAME, specs[s.idx - 1].cout))
            s.cin_pad = 16 if s.sources[0][1] == L.SRC_INPUT else sum(c for _, _, c in s.sources)   # stem: 32 when the input is repacked
            if s.sources[0][1] == L.SRC_INPUT and s.cin_real > 16:
                raise L.NunetError("input_channels > 16 is not supported by the stem loader")
            if s.sources[0][1] != L.SRC_INPUT and s.cin_pad != s.cin_real:
                raise L.NunetError("channel bookkeeping mismatch in %s" % node)
            specs.append(s)
        node_out[node] = specs[-1].idx
    for s in specs:
        off = 0
        for j, (pi, kind, c) in enumerate(s.sources):
            if pi >= 0:
                specs[pi].consumers.append((s.idx, off, kind, j))
            off += c
    for head_name, node in model._heads:
        specs[node_out[node]].head = getattr(model, head_name)
    return specs


class _Streams(object):
    """One compute stream, one weight-gradient stream and one BN scratch per resolution level."""

    def __init__(self, device):
        self.device = device
        self.lvl = [torch.cuda.Stream(device=device) for _ in NB_FILTER]
        self.wg = [torch.cuda.Stream(device=device) for _ in NB_FILTER]
        self.ws = [ops.BnWorkspace(device) for _ in NB_FILTER]

    def all(self, with_wg):
        return self.lvl + (self.wg if with_wg else [])

    def fork(self, main, with_wg=False):
        for st_ in self.all(with_wg):
            st_.wait_stream(main)

    def join(self, main, with_wg=False):
        for st_ in self.all(with_wg):
            main.wait_stream(st_)


class _NoCtx(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NOCTX = _NoCtx()


def _tma_width(w):
    """The stacked Cout = 32 kernel and the TMA weight-gradient kernel take rows of up to 254 pixels (box dimension limit 256 incl.
    padding)."""
    return w + 2 <= 256


STRIPS = _os.environ.get("NUNET_NO_STRIPS") is None


def _mat_width(w):
    """Operands are materialised (and fed by tensor loads) at this row width: always, since the forward / dgrad kernel describes wider
    rows as column strips (nunet_conv_plan_t.strips); NUNET_NO_STRIPS=1 restores the transforming-loader path for rows > 254 px."""
    return STRIPS or _tma_width(w)


def _stacked(s, N, h, w, cin_pad):
    """True if this layer's forward conv runs on the stacked Cout = 32 kernel (conv_c32.cu): all its sources are materialised
    activations at this resolution (always the case where the TMA width limit holds) and the shape is eligible."""
    if s.cout != 32 or not _mat_width(w) or not (MATERIALIZE_NODE_OUTPUTS and MATERIALIZE_UPSAMPLED and MATERIALIZE_ALL and MATERIALIZE_POOLED):
        return False
    return ops.c32_eligible(N, h, w, cin_pad, s.cout)


def _cin_pad(s, packed_input):
    return 32 if (packed_input and s.sources[0][1] == L.SRC_INPUT) else s.cin_pad


_WEIGHTS_EPOCH = 0   # bumped by the fused optimizers (optim.py), which update the fp32 masters through raw pointers (no autograd version bump)


def weights_updated():
    global _WEIGHTS_EPOCH
    _WEIGHTS_EPOCH += 1


_CAPTURING_TRAIN = False   # set while a training-step graph is being captured (fork/join of the level streams is capturable)


def _use_level_streams(train):
    return LEVEL_STREAMS and train and (_CAPTURING_TRAIN or not torch.cuda.is_current_stream_capturing())


class _State(object):
    """Per-model engine state: layer specs, packed-weight cache, BN scratch."""

    def __init__(self, model):
        self.specs = build_specs(model)
        self.packs = {}
        self.ws = None
        self.device = None

    def workspace(self, device):
        if self.ws is None or self.device != device:
            self.ws = ops.BnWorkspace(device)
            self.device = device
            self.packs = {}
            # one self-resetting last-CTA counter per layer and direction (BatchNorm finalisation in the producing kernel's tail)
            self.counters = torch.zeros(2 * len(self.specs), dtype=torch.int32, device=device)
        return self.ws

    def counter(self, spec, backward=False):
        i = 2 * spec.idx + (1 if backward else 0)
        return self.counters[i:i + 1]

    def pack(self, spec, plan, mode):
        w = spec.conv.weight
        key = (spec.idx, mode)
        tag = (w._version, w.data_ptr(), plan.NT, plan.CK, _WEIGHTS_EPOCH)
        hit = self.packs.get(key)
        if hit is not None and (hit[0] == tag or (_CAPTURING_TRAIN and hit[0][1:4] == tag[1:4])):
            return hit[1]
        out = hit[1] if (hit is not None and hit[1].numel() == plan.wpack_elems and hit[1].device == w.device) else None
        out = ops.pack_weights(w, mode, plan, out)
        self.packs[key] = (tag, out)
        return out


    def streams(self, device):
        ms = self.__dict__.get("_ms")
        if ms is None or ms.device != device:
            ms = self.__dict__["_ms"] = _Streams(device)
        return ms

    def pack_all(self, N, H, W, want_dgrad, packed_input):
        """Re-packs every stale weight image of the step (forward and dgrad) with one batched launch."""
        items, keys = [], []
        for s in self.specs:
            h, w = H >> s.level, W >> s.level
            cinp = _cin_pad(s, packed_input)
            modes = [(L.PACK_FWD_C32 if _stacked(s, N, h, w, cinp) else L.PACK_FWD, ops.conv_plan(N, h, w, cinp, s.cout))]
            if want_dgrad and s.sources[0][0] >= 0:
                dstack = (s.which == 2 and FUSE_BLOCK_ROUTER and s.cin_pad == 32 and _mat_width(w) and ops.c32_eligible(N, h, w, s.cout, s.cin_pad))
                modes.append((L.PACK_DGRAD_C32 if dstack else L.PACK_DGRAD, ops.conv_plan(N, h, w, s.cout, s.cin_pad)))
            wt = s.conv.weight
            if wt.dtype != torch.float32 or not wt.is_contiguous():
                continue   # the per-layer path converts
            for mode, plan in modes:
                key = (s.idx, mode)
                tag = (wt._version, wt.data_ptr(), plan.NT, plan.CK, _WEIGHTS_EPOCH)
                hit = self.packs.get(key)
                if hit is not None and hit[0] == tag and not _CAPTURING_TRAIN:   # a captured step always re-packs (replayed after every update)
                    continue
                out = hit[1] if (hit is not None and hit[1].numel() == plan.wpack_elems and hit[1].device == wt.device) else \
                    torch.empty(plan.wpack_elems, dtype=torch.bfloat16, device=wt.device)
                items.append((wt.detach(), mode, plan, out))
                keys.append((key, tag, out))
        if items:
            ops.pack_weights_batched(items)
            for key, tag, out in keys:
                self.packs[key] = (tag, out)


def _state(model):
    st = model.__dict__.get("_nunet_state")
    if st is None:
        st = _State(model)
        model.__dict__["_nunet_state"] = st
    return st


# SyncBN / data-parallel hooks (set by dist.DataParallel): callables that all-reduce an fp64 [2,C] tensor in place.
def _world(model):
    return model.__dict__.get("_nunet_dp")


class _Record(object):
    __slots__ = ("y", "a", "a_up", "a_pool", "scale", "shift", "mean", "invstd", "h", "w", "logits")


def _sources(spec, recs, x, x32=None):
    """Loader descriptors of a conv: node outputs are read from their materialised activation (copy / pool / upsample
    loaders without BN math), a block's first conv output through the fused BN+ReLU loader."""
    srcs = []
    for pi, kind, c in spec.sources:
        if pi < 0:
            srcs.append(ops.Source(L.SRC_ACT, x32, 32) if x32 is not None else ops.Source(kind, x, c))
        else:
            r = recs[pi]
            if kind == L.SRC_UP and r.a_up is not None:
                srcs.append(ops.Source(L.SRC_ACT, r.a_up, c))
            elif kind == L.SRC_POOL and r.a_pool is not None:
                srcs.append(ops.Source(L.SRC_ACT, r.a_pool, c))
            elif r.a is not None:
                srcs.append(ops.Source(L.ACT_KIND[kind], r.a, c))
            else:
                srcs.append(ops.Source(kind, r.y, c, r.scale, r.shift))
    return srcs


def _forward_layer(s, recs, x, x32, N, H, W, train, dp, ws, dev, st):
    """One conv + BatchNorm (+ the materialised operands its consumers read) on the current stream."""
    h, w = H >> s.level, W >> s.level
    srcs = _sources(s, recs, x, x32)
    cinp = _cin_pad(s, x32 is not None)
    plan = ops.conv_plan(N, h, w, cinp, s.cout)
    stacked = _stacked(s, N, h, w, cinp) and all(sr.kind == L.SRC_ACT for sr in srcs)
    wp = st.pack(s, plan, L.PACK_FWD_C32 if stacked else L.PACK_FWD)
    r = _Record()
    r.h, r.w, r.a, r.a_up, r.a_pool = h, w, None, None, None
    r.scale = torch.empty(s.cout, dtype=torch.float32, device=dev)
    r.shift = torch.empty(s.cout, dtype=torch.float32, device=dev)
    bn = s.bn
    if train:
        # train-mode BatchNorm statistics come out of the conv epilogue and are finalised by the conv's last CTA (csrc/bn_tail.cuh)
        r.mean = torch.empty(s.cout, dtype=torch.float32, device=dev)
        r.invstd = torch.empty(s.cout, dtype=torch.float32, device=dev)
        momentum = bn.momentum if bn.momentum is not None else 1.0 / float(int(bn.num_batches_tracked) + 1)
        track = bn.track_running_stats and bn.running_mean is not None
        rm = bn.running_mean if track else None
        rv = bn.running_var if track else None
        count = float(N * h * w)
        if dp is None or not dp.sync_bn:
            tail = ops.bn_tail(st.counter(s), count, bn.weight, bn.bias, s.conv.bias, rm, rv, momentum, bn.eps, r.scale, r.shift, r.mean, r.invstd)
            r.y, _ = ops.conv3x3(N, h, w, srcs, s.cout, wp, want_stats=True, stacked=stacked, tail=tail)
        elif dp.exchange is not None:
            # SyncBN inside the conv's tail: the last CTAs of all ranks exchange their sums through peer memory (dist.PeerExchange)
            tail = ops.bn_tail(st.counter(s), count * dp.world_size, bn.weight, bn.bias, s.conv.bias, rm, rv, momentum, bn.eps, r.scale, r.shift,
                               r.mean, r.invstd, xch=dp.exchange.slot(2 * s.idx))
            r.y, _ = ops.conv3x3(N, h, w, srcs, s.cout, wp, want_stats=True, stacked=stacked, tail=tail)
        else:
            sums = torch.empty((2, s.cout), dtype=torch.float64, device=dev)
            tail = ops.bn_tail(st.counter(s), count, None, None, None, None, None, momentum, bn.eps, None, None, None, None, sums_out=sums)
            r.y, _ = ops.conv3x3(N, h, w, srcs, s.cout, wp, want_stats=True, stacked=stacked, tail=tail)
            dp.all_reduce_sum(sums)
            ops.bn_finalize_from_sums(sums, count * dp.world_size, bn.weight, bn.bias, s.conv.bias, rm, rv, momentum, bn.eps,
                                      r.scale, r.shift, r.mean, r.invstd)
    else:
        r.mean = r.invstd = None
        r.y, _ = ops.conv3x3(N, h, w, srcs, s.cout, wp, want_stats=False, stacked=stacked)
        ops.bn_fold_eval(bn.weight, bn.bias, s.conv.bias, bn.running_mean, bn.running_var, bn.eps, r.scale, r.shift)
    y = r.y
    # One "activate" pass over y writes whatever the consumers read: the activation itself (same-resolution consumers), its 2x2
    # max-pool, its bilinear 2x upsample, and the logits of a 1x1 head.  (A block's first conv output and the pooled operand are
    # only worth materialising where the TMA-fed kernels take them.)
    r.logits = None
    kinds = [c[2] for c in s.consumers]
    mat = (s.which == 2 or (MATERIALIZE_ALL and _mat_width(w))) and MATERIALIZE_NODE_OUTPUTS and bool(kinds)
    want_up = mat and MATERIALIZE_UPSAMPLED and L.SRC_UP in kinds
    want_pool = mat and MATERIALIZE_POOLED and _mat_width(w // 2) and L.SRC_POOL in kinds
    want_a = mat and (L.SRC_SAME in kinds or (L.SRC_UP in kinds and not want_up) or (L.SRC_POOL in kinds and not want_pool))
    head = None
    if s.head is not None and s.cout == 32 and s.head.out_channels <= 8 and not (want_up or want_pool):
        K = s.head.out_channels
        head = (s.head.weight.detach().view(K, -1), s.head.bias.detach() if s.head.bias is not None else None)
    if want_a or want_up or want_pool or head is not None:
        r.a, r.a_pool, r.a_up, r.logits = ops.bn_activate(y, r.scale, r.shift, want_a, want_pool, want_up, head)
    return r


def _forward(model, x, train, want_dgrad=False):
    st = _state(model)
    specs = st.specs
    if x.dim() != 4:
        raise ValueError("expected NCHW input")
    ops.require_cuda(x, "input")
    # Accepted inputs: the reference's fp32 NCHW tensor (trains.py:113-117); and, as the device-side input pipeline (ops.input_u8),
    # a raw uint8 NHWC image batch (normalised here with the loader's arithmetic) or its already packed bf16 NHWC32 form.
    x32 = None
    if x.dtype == torch.uint8:
        N, H, W, Cin = x.shape
        x = x.contiguous()
    elif x.dtype == torch.bfloat16 and x.shape[-1] == 32 and x.is_contiguous():
        N, H, W, _ = x.shape
        Cin, x32 = specs[0].cin_real, x
    else:
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        N, Cin, H, W = x.shape
    if Cin != specs[0].cin_real:
        raise ValueError("expected %d input channels, got %d" % (specs[0].cin_real, Cin))
    if H % 16 or W % 16:
        raise ValueError("input height/width must be multiples of 16 (four 2x2 poolings), got %dx%d" % (H, W))
    dev = x.device
    ws = st.workspace(dev)
    dp = _world(model)
    recs = []
    packed_input = PACK_INPUT and _mat_width(W)
    st.pack_all(N, H, W, want_dgrad, packed_input)
    norm = model.__dict__.get("_nunet_input_norm") or {}
    if x32 is not None:
        if not packed_input:
            raise L.NunetError("pre-packed bf16 NHWC32 input needs rows of at most 254 pixels")
    elif x.dtype == torch.uint8:
        if packed_input:
            x32 = ops.input_u8(x, nhwc32=True, **norm)
        else:
            x = ops.input_u8(x, nhwc32=False, **norm)
    elif packed_input:
        x32 = ops.input_to_nhwc32(x)
    ms = st.streams(dev) if _use_level_streams(train) else None
    main = torch.cuda.current_stream(dev)
    done = {}   # layer idx of a node output -> event recorded after the node's last forward launch
    if ms is not None:
        ms.fork(main)
    for s in specs:
        lvl_stream = ms.lvl[s.level] if ms is not None else None
        with (torch.cuda.stream(lvl_stream) if ms is not None else _NOCTX):
            if ms is not None and s.which == 1:
                for pi, _, _ in s.sources:
                    if pi >= 0 and specs[pi].level != s.level:
                        lvl_stream.wait_event(done[pi])
            recs.append(_forward_layer(s, recs, x, x32, N, H, W, train, dp, ms.ws[s.level] if ms is not None else ws, dev, st))
            if ms is not None and s.which == 2:
                done[s.idx] = torch.cuda.Event()
                done[s.idx].record(lvl_stream)
    if train:
        nbt = [s.bn.num_batches_tracked for s in specs if s.bn.num_batches_tracked is not None]
        if nbt:
            torch._foreach_add_(nbt, 1)
    logits = []
    for s in specs:
        if s.head is not None:
            r = recs[s.idx]
            K = s.head.out_channels
            if r.logits is not None:      # came out of the node's activate pass
                logits.append((s.idx, r.logits))
                continue
            with (torch.cuda.stream(ms.lvl[s.level]) if ms is not None else _NOCTX):
                logits.append((s.idx, ops.head_fwd(r.y, r.scale, r.shift, s.head.weight.detach().view(K, -1), s.head.bias.detach(), N, r.h, r.w)))
    if ms is not None:
        ms.join(main)
    return (x, x32, (N, H, W)), recs, logits


def _backward(model, xs, recs, grad_logits, grad_alloc=None):
    """grad_logits: {layer idx: dlogits fp32 NCHW}. Returns {id(param): grad}.  grad_alloc(p): where p's gradient is written
    (captured steps: views of one flat buffer); default: a fresh tensor, or the data-parallel bucket view."""
    st = _state(model)
    specs = st.specs
    x, x32, (N, H, W) = xs
    dev = x.device
    ws = st.workspace(dev)
    dp = _world(model)
    grads, dxs, pre = {}, {}, {}   # dxs: layer -> per-source dense dgrad outputs; pre: layer -> g_z and BN-backward coefficients from a fused router
    main = torch.cuda.current_stream(dev)
    ms = st.streams(dev) if _use_level_streams(True) else None
    side = None
    if ms is None and WGRAD_SIDE_STREAM and not torch.cuda.is_current_stream_capturing():
        side = st.__dict__.get("side_stream")
        if side is None:
            side = st.__dict__["side_stream"] = torch.cuda.Stream(device=dev)
    # captured steps write into the slot's flat buffer; its single all-reduce is issued after the replay (_NetFunction.backward)
    buckets = dp.buckets if (dp is not None and grad_alloc is None) else None
    if buckets is not None:
        buckets.begin(dev)
        buckets.side = ms.all(True) if ms is not None else side
    keep = []      # level streams: every cross-stream tensor of the pass stays referenced until the final join is enqueued
    dx_done = {}   # layer idx -> event recorded after its dgrad
    if ms is not None:
        ms.fork(main, with_wg=True)

    def galloc(p):
        if grad_alloc is not None:
            return grad_alloc(p)
        return buckets.view(p) if buckets is not None else torch.empty(p.shape, dtype=torch.float32, device=dev)

    def gdone(p, g):
        grads[id(p)] = g
        if buckets is not None:
            buckets.ready(p)

    for s in reversed(specs):
        r = recs[s.idx]
        h, w = r.h, r.w
        S = ms.lvl[s.level] if ms is not None else None
        ready = []   # (param, grad) finished by this layer; announced on the caller's stream (bucket all-reduce launches there)
        with (torch.cuda.stream(S) if ms is not None else _NOCTX):
            if ms is not None:
                for c_ in s.consumers:
                    if specs[c_[0]].level != s.level:
                        S.wait_event(dx_done[c_[0]])
            fused = pre.pop(s.idx, None)   # this layer's router ran inside the epilogue of its consumer's dgrad
            head_grads = None
            cons = []
            for (m, off, kind, j) in (s.consumers if fused is None else ()):
                dxm = dxs[m][j]            # the consumer's dgrad wrote our channels as one dense tensor
                if kind == L.SRC_UP:   # transposed bilinear gather in its own kernel; the router then sees a dense SAME consumer
                    cons.append(ops.Consumer(ops.up_transpose(dxm, s.cout, 0, N, h, w, s.cout), s.cout, 0, L.SRC_SAME))
                else:
                    cons.append(ops.Consumer(dxm, s.cout, 0, kind))
            dl = wh = None
            if s.head is not None:
                dl = grad_logits.get(s.idx)
                if dl is not None:
                    K = s.head.out_channels
                    wh = s.head.weight.detach().view(K, -1)
                    dwh, dbh = galloc(s.head.weight), galloc(s.head.bias)
                    # one class (the reference's default): the head's own gradients are accumulated by the router pass itself
                    if K == 1 and s.cout == 32 and (dp is None or not dp.sync_bn or dp.exchange is not None) and all(c[2] == L.SRC_SAME for c in s.consumers):
                        head_grads = (dwh, dbh)
                    else:
                        ops.head_bwd_params(r.y, r.scale, r.shift, dl, N, h, w, K, dwh, dbh)
                    ready += [(s.head.weight, dwh), (s.head.bias, dbh)]
            if not cons and dl is None and fused is None:
                dl = torch.zeros((N, 1, h, w), dtype=torch.float32, device=dev)
                wh = torch.zeros((1, s.cout), dtype=torch.float32, device=dev)
            count = float(N * h * w)
            if fused is not None:
                gz, dgamma, dbeta, k1, k2, k3, sums = fused
                if sums is not None:     # SyncBN: the fused router only produced this rank's sums
                    dp.all_reduce_sum(sums)
                    ops.bn_backward_from_sums(sums, count * dp.world_size, s.bn.weight, r.mean, r.invstd, dgamma, dbeta, k1, k2, k3,
                                              grad_scale=1.0 / dp.world_size)
            else:
                dgamma, dbeta = galloc(s.bn.weight), galloc(s.bn.bias)
                k1, k2, k3 = (torch.empty(s.cout, dtype=torch.float32, device=dev) for _ in range(3))
            if fused is not None:
                pass
            elif dp is None or not dp.sync_bn or dp.exchange is not None:
                # the router's last CTA reduces its partial sums and emits dgamma / dbeta / k1..k3 (csrc/bn_tail.cuh); under SyncBN the
                # last CTAs of all ranks first exchange their sums through peer memory.  Local gradients are those of the LOCAL mean
                # loss (= world x the global-loss gradient): dgamma / dbeta are scaled by 1/world so the later AVG keeps the true value
                wsz = dp.world_size if (dp is not None and dp.sync_bn) else 1
                tail = ops.bn_bwd_tail(st.counter(s, True), count * wsz, s.bn.weight, r.mean, r.invstd, dgamma, dbeta, k1, k2, k3, grad_scale=1.0 / wsz,
                                       xch=dp.exchange.slot(2 * s.idx + 1) if wsz > 1 else None)
                gz, _ = ops.grad_router(N, h, w, s.cout, r.y, r.scale, r.shift, r.mean, r.invstd, cons, dlogits=dl, wh=wh, tail=tail,
                                        head_grads=head_grads)
            else:
                sums = torch.empty((2, s.cout), dtype=torch.float64, device=dev)
                tail = ops.bn_bwd_tail(st.counter(s, True), count, None, None, None, None, None, None, None, None, sums_out=sums)
                gz, _ = ops.grad_router(N, h, w, s.cout, r.y, r.scale, r.shift, r.mean, r.invstd, cons, dlogits=dl, wh=wh, tail=tail)
                dp.all_reduce_sum(sums)
                # local gradients are those of the LOCAL mean loss (= world_size x the global-loss gradient); the all-reduced sums
                # therefore give world_size x dgamma/dbeta on every rank: scale by 1/world so the later AVG keeps the true value
                ops.bn_backward_from_sums(sums, count * dp.world_size, s.bn.weight, r.mean, r.invstd, dgamma, dbeta, k1, k2, k3,
                                          grad_scale=1.0 / dp.world_size)
            ready += [(s.bn.weight, dgamma), (s.bn.bias, dbeta)]
            gy = ops.bn_bwd_apply(gz, r.y, k1, k2, k3, out=gz)  # in place: g_z is dead once the BN sums are reduced
            srcs = _sources(s, recs, x, x32)
            dw = galloc(s.conv.weight)
            creal = s.cin_real if (s.sources[0][0] < 0 and x32 is not None) else 0
            if ms is not None:
                # weight gradients are only needed when backward returns: each level has a second stream for them
                gy_ready = torch.cuda.Event()
                gy_ready.record(S)
                ms.wg[s.level].wait_event(gy_ready)
                with torch.cuda.stream(ms.wg[s.level]):
                    ops.conv3x3_wgrad(N, h, w, srcs, s.cout, gy, dw, cin_real=creal)
                keep.append(gy)
            elif side is None:
                ops.conv3x3_wgrad(N, h, w, srcs, s.cout, gy, dw, cin_real=creal)
            else:
                # run it on a second stream so that the HBM-bound router / BN kernels of the following layers execute while
                # it occupies the tensor pipes
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    ops.conv3x3_wgrad(N, h, w, srcs, s.cout, gy, dw, cin_real=creal)
                gy.record_stream(side)
                dw.record_stream(side)
            ready.append((s.conv.weight, dw))
            if s.conv.bias is not None:
                # train-mode BatchNorm subtracts the batch mean: the conv bias has an exactly zero gradient
                ready.append((s.conv.bias, galloc(s.conv.bias) if (buckets is not None or grad_alloc is not None) else torch.zeros_like(s.conv.bias)))
            if s.sources[0][0] >= 0:
                plan = ops.conv_plan(N, h, w, s.cout, s.cin_pad)
                # a dgrad that produces 32 channels (level-0 second convs) runs on the stacked kernel too
                dstack = (s.which == 2 and FUSE_BLOCK_ROUTER and s.cin_pad == 32 and _mat_width(w) and ops.c32_eligible(N, h, w, s.cout, s.cin_pad))
                wpd = st.pack(s, plan, L.PACK_DGRAD_C32 if dstack else L.PACK_DGRAD)
                gsrc = [ops.Source(L.SRC_ACT, gy, s.cout)]
                if s.which == 2 and FUSE_BLOCK_ROUTER:
                    # the block's first conv has exactly one consumer (this conv): its gradient router — ReLU mask, g_z, the two
                    # BatchNorm-backward sums and (single GPU) their finalisation — runs in this dgrad's epilogue
                    sp, rp = specs[s.idx - 1], recs[s.idx - 1]
                    dg1, db1 = galloc(sp.bn.weight), galloc(sp.bn.bias)
                    kk = [torch.empty(sp.cout, dtype=torch.float32, device=dev) for _ in range(3)]
                    msk = ops.Mask(rp.y, rp.scale, rp.shift, rp.mean, rp.invstd)
                    if dp is None or not dp.sync_bn or dp.exchange is not None:
                        wsz = dp.world_size if (dp is not None and dp.sync_bn) else 1
                        bt, sums1 = ops.bn_bwd_tail(st.counter(sp, True), count * wsz, sp.bn.weight, rp.mean, rp.invstd, dg1, db1, kk[0], kk[1], kk[2],
                                                    grad_scale=1.0 / wsz, xch=dp.exchange.slot(2 * sp.idx + 1) if wsz > 1 else None), None
                    else:
                        sums1 = torch.empty((2, sp.cout), dtype=torch.float64, device=dev)
                        bt = ops.bn_bwd_tail(st.counter(sp, True), count, None, None, None, None, None, None, None, None, sums_out=sums1)
                    gz1, _ = ops.conv3x3(N, h, w, gsrc, s.cin_pad, wpd, cls="conv_dgrad", mask=msk, bwd_tail=bt, stacked=dstack)
                    pre[sp.idx] = (gz1, dg1, db1, kk[0], kk[1], kk[2], sums1)
                else:
                    dxs[s.idx], _ = ops.conv3x3(N, h, w, gsrc, s.cin_pad, wpd, cls="conv_dgrad", out_segs=[c for _, _, c in s.sources])
                if ms is not None:
                    dx_done[s.idx] = torch.cuda.Event()
                    dx_done[s.idx].record(S)
        for p_, g_ in ready:
            gdone(p_, g_)
        if ms is None:
            for c_ in s.consumers:  # a consumer's dX is dead once its last (earliest) producer has routed it
                if specs[c_[0]].sources[0][0] == s.idx:
                    dxs.pop(c_[0], None)
    if ms is not None:
        ms.join(main, with_wg=True)
    if side is not None:
        main.wait_stream(side)
    if buckets is not None:
        buckets.finish()
    del keep
    return grads


# ------------------------------------------------------------------------------------------------ captured training steps
# A training step is ~380 launches on ten streams; at small batch (and on the host-synchronised end-to-end path) enqueueing
# them costs more than running them.  After one eager step per (input shape, parameter set), forward and backward are each
# captured into a CUDA graph (same kernels, same per-stream order: bitwise identical results) and replayed: two graph
# launches per step.  The optimizer stays outside (it updates the fp32 masters in place; the graph re-packs them).
GRAPH_TRAIN = _os.environ.get("NUNET_GRAPH_TRAIN", "1") == "1"
# Capturing the data-parallel step (NCCL all-reduces inside the graphs) is wired but OFF: with torch 2.11 / NCCL 2.28 the
# capture of the bucketed async all-reduce + SyncBN exchanges deadlocked on 2 GPUs (gpurun, round 1); DP steps run eagerly.
GRAPH_TRAIN_DP = _os.environ.get("NUNET_GRAPH_TRAIN_DP", "0") == "1"


class _TrainGraph(object):
    __slots__ = ("seen", "pool", "x", "fwd", "outs", "recs", "xs", "head_layers", "launches_f", "bwd", "gl", "flat", "views", "launches_b",
                 "packs", "failed", "gen", "bwd_gen", "live")

    def __init__(self):
        self.seen, self.fwd, self.bwd, self.failed, self.packs = 0, None, None, False, {}
        self.gen = self.bwd_gen = 0   # forward replays / the last one whose backward ran
        self.live = None              # weak reference to the autograd node of the last captured forward


def _train_graph(model, x, params):
    """The graph slot of this (shape, parameter set), or None when the step has to run eagerly."""
    if not GRAPH_TRAIN or ops.kernel_timing_enabled():
        return None
    dp_ = _world(model)
    if dp_ is not None and not GRAPH_TRAIN_DP and not (dp_.sync_bn and dp_.exchange is not None or not dp_.sync_bn):
        return None      # SyncBN through NCCL all-reduces: 60 collectives inside the step, runs eagerly
    if not (torch.is_tensor(x) and x.is_cuda and x.dim() == 4 and x.dtype in (torch.float32, torch.uint8, torch.bfloat16) and x.is_contiguous()):
        return None
    if torch.cuda.is_current_stream_capturing():
        return None
    st = _state(model)
    if any(s.bn.momentum is None for s in st.specs):      # cumulative moving average: the factor changes every step
        return None
    graphs = st.__dict__.setdefault("train_graphs", {})
    key = (tuple(x.shape), x.dtype, x.device.index, tuple(p.data_ptr() for p in params), tuple(p.requires_grad for p in params))
    tg = graphs.get(key)
    if tg is None:
        if len(graphs) >= 4:                               # each slot pins one step's activations
            graphs.pop(next(iter(graphs)))
        tg = graphs[key] = _TrainGraph()
    tg.seen += 1
    if tg.seen <= 1 or tg.failed:
        return None
    if tg.gen != tg.bwd_gen:      # the previous captured forward still waits for its backward: its activations must survive
        if tg.live is not None and tg.live() is None:
            tg.bwd_gen = tg.gen   # ... unless its outputs were dropped without a backward (the autograd node is gone)
        else:
            return None
    return tg


def _capture(tg, model, fn):
    """Runs fn() under stream capture with the slot's private memory pool and packed-weight cache."""
    global _CAPTURING_TRAIN
    st = _state(model)
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    saved_packs, st.packs = st.packs, tg.packs
    _CAPTURING_TRAIN = True
    n0 = ops.launch_count()
    try:
        with torch.cuda.graph(g, pool=tg.pool, capture_error_mode="thread_local"):
            res = fn()
    finally:
        _CAPTURING_TRAIN = False
        st.packs = saved_packs
    launches = ops.launch_count() - n0
    ops.add_launches(-launches)       # nothing ran yet: replays are what gets counted
    return g, res, launches


class _NetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, *params):
        tg = _train_graph(model, x, params)
        ctx.model, ctx.params, ctx.tg = model, params, tg
        if tg is None:
            x, recs, logits = _forward(model, x, True, want_dgrad=True)
            ctx.x, ctx.recs = x, recs
            ctx.head_layers = [i for i, _ in logits]
            return tuple(t for _, t in logits)
        if tg.fwd is None:
            try:
                tg.pool = torch.cuda.graph_pool_handle()
                tg.x = x.clone()
                tg.fwd, (tg.xs, tg.recs, logits), tg.launches_f = _capture(tg, model, lambda: _forward(model, tg.x, True, want_dgrad=True))
                tg.head_layers = [i for i, _ in logits]
                tg.outs = tuple(t for _, t in logits)
            except Exception as exc:      # capture is an optimisation: fall back to the eager step for this shape
                import warnings
                warnings.warn("nunet: CUDA-graph capture of the training forward failed (%s); running eagerly" % (exc,))
                tg.failed, tg.fwd, ctx.tg = True, None, None
                x, recs, logits = _forward(model, x, True, want_dgrad=True)
                ctx.x, ctx.recs = x, recs
                ctx.head_layers = [i for i, _ in logits]
                return tuple(t for _, t in logits)
        tg.x.copy_(x)
        tg.fwd.replay()
        ops.add_launches(tg.launches_f)
        tg.gen += 1
        ctx.gen = tg.gen
        try:
            tg.live = _weakref.ref(ctx)
        except TypeError:
            tg.live = None
        ctx.head_layers = tg.head_layers
        return tuple(t.clone() for t in tg.outs)

    @staticmethod
    def backward(ctx, *grad_logits):
        tg = ctx.tg
        if tg is None:
            gl = {}
            for i, g in zip(ctx.head_layers, grad_logits):
                if g is not None:
                    gl[i] = g.contiguous().float()
            grads = _backward(ctx.model, ctx.x, ctx.recs, gl)
            ctx.recs = None
            return (None, None) + tuple(grads.get(id(p)) if p.requires_grad else None for p in ctx.params)
        if tg.bwd is None:
            dev = tg.x.device
            st = _state(ctx.model)
            tg.gl = {i: torch.zeros_like(o) for i, o in zip(tg.head_layers, tg.outs)}
            offs, total = {}, 0
            for p in ctx.params:
                offs[id(p)] = (total, p.numel(), tuple(p.shape))
                total += (p.numel() + 31) // 32 * 32
            tg.flat = torch.zeros(total, dtype=torch.float32, device=dev)

            def alloc(p):
                o, n, shape = offs[id(p)]
                return tg.flat[o:o + n].view(shape)
            tg.views = offs
            if tg.failed:
                pass
            elif _world(ctx.model) is None or not GRAPH_TRAIN_DP:
                try:
                    tg.bwd, _, tg.launches_b = _capture(tg, ctx.model, lambda: _backward(ctx.model, tg.xs, tg.recs, tg.gl, grad_alloc=alloc))
                except Exception as exc:
                    import warnings
                    warnings.warn("nunet: CUDA-graph capture of the training backward failed (%s); running eagerly" % (exc,))
                    tg.failed, tg.bwd = True, None
            else:
                # data parallel: backward writes into the bucket buffer it all-reduces in place (captured NCCL launches);
                # the reduced gradients are then gathered into the slot's flat buffer inside the same graph
                def bwd_dp():
                    grads = _backward(ctx.model, tg.xs, tg.recs, tg.gl)
                    for p in ctx.params:
                        g = grads.get(id(p))
                        if g is not None:
                            alloc(p).copy_(g)
                tg.bwd, _, tg.launches_b = _capture(tg, ctx.model, bwd_dp)
        for i, g in zip(tg.head_layers, grad_logits):
            if g is None:
                tg.gl[i].zero_()
            else:
                tg.gl[i].copy_(g)
        if tg.bwd is not None:
            tg.bwd.replay()
            ops.add_launches(tg.launches_b)
            tg.bwd_gen = ctx.gen
            dp_ = _world(ctx.model)
            if dp_ is not None and not GRAPH_TRAIN_DP:
                # data parallel: ONE all-reduce (AVG) of the flat gradient buffer over NCCL / NVLink right behind the captured backward
                import torch.distributed as _dist
                _dist.all_reduce(tg.flat, op=_dist.ReduceOp.AVG if _dist.get_backend(dp_.group) == "nccl" else _dist.ReduceOp.SUM, group=dp_.group)
                if _dist.get_backend(dp_.group) != "nccl":
                    tg.flat.div_(dp_.world_size)
        else:                         # backward capture failed earlier: eager backward on the captured forward's activations
            grads = _backward(ctx.model, tg.xs, tg.recs, tg.gl)
            tg.bwd_gen = ctx.gen
            return (None, None) + tuple(grads.get(id(p)) if p.requires_grad else None for p in ctx.params)
        # The graph's flat gradient buffer is rewritten by the next replay.  When no parameter holds a gradient (the usual
        # optimizer.zero_grad() before backward) autograd adopts the returned tensors as p.grad, and the next replay can only happen
        # after the next zero_grad released them: hand out views of the buffer itself — stable addresses for the fused optimizer's
        # pointer table, no 36 MB copy.  Otherwise (gradient accumulation) the caller gets a private copy.
        flat = tg.flat if all(p.grad is None for p in ctx.params) else tg.flat.clone()
        out = []
        for p in ctx.params:
            if p.requires_grad:
                o, n, shape = tg.views[id(p)]
                out.append(flat[o:o + n].view(shape))
            else:
                out.append(None)
        return (None, None) + tuple(out)


# Eval-mode forwards (val.py:87-98) are ~65 short launches; at small batch the step is launch-bound, so the whole forward is
# captured once per input shape into a CUDA graph and replayed (re-captured when any parameter changes).
GRAPH_EVAL = _os.environ.get("NUNET_GRAPH_EVAL", "1") == "1"


class _EvalGraph(object):
    __slots__ = ("tag", "graph", "x", "outs", "packs")

MAX_EVAL_GRAPHS = 8      # each slot pins one forward's activations and its own packed-weight images


def _weights_tag(params):
    return (params[0].data_ptr(), sum(p._version for p in params), _WEIGHTS_EPOCH)


def _graphed_eval(model, x, params):
    st = _state(model)
    graphs = st.__dict__.setdefault("graphs", {})
    key = (tuple(x.shape), x.dtype, x.device.index)
    tag = _weights_tag(params)
    g = graphs.get(key)
    if g is None or g.tag != tag:
        # A captured graph holds raw pointers to the packed weight images it was captured with.  The shared cache (st.packs) re-packs an
        # image IN PLACE when another input shape needs a different tiling (CK / NT), which would silently corrupt a graph of this
        # shape replayed later: every slot owns its images, like the training slots do.
        graphs.pop(key, None)
        while len(graphs) >= MAX_EVAL_GRAPHS:
            graphs.pop(next(iter(graphs)))
        g = _EvalGraph()
        g.packs = {}
        st.workspace(x.device)       # (re)initialise the per-device state BEFORE the slot's cache is swapped in: it resets st.packs
        saved_packs, st.packs = st.packs, g.packs
        try:
            with torch.no_grad():
                _forward(model, x, False)                  # eager warm-up: fills this slot's packed-weight cache, loads kernels
            torch.cuda.current_stream().synchronize()
            g.tag = tag
            g.x = x.clone()
            g.graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(g.graph):
                _, _, logits = _forward(model, g.x, False)
        finally:
            st.packs = saved_packs
        g.outs = tuple(t for _, t in logits)
        graphs[key] = g
    g.x.copy_(x)
    g.graph.replay()
    return tuple(t.clone() for t in g.outs)


# ------------------------------------------------------------------------------------------------ fp32 check mode
# Evaluation of the same graph, forward AND backward, in CUDA-core fp32 with fp64 reductions (csrc/check_fp32.cu): the
# north_star's "<= 1e-4 in an fp32 check mode" bound against the reference's logits, and the reference's gradients to <= 1e-3
# rel-L2 (tests/test_model_gpu.py).  Selected explicitly; never a fallback of the product path.
CHECK_FP32 = _os.environ.get("NUNET_FP32_CHECK", "0") == "1"
_ROUTE_MODE = {L.SRC_SAME: 0, L.SRC_POOL: 1, L.SRC_UP: 2}


def _check_call(name, *args):
    ops._call("check_fp32", 1, None, name, getattr(L.lib(), name), *args)


class _CheckRecord(object):
    __slots__ = ("buf", "cin", "y", "a", "scale", "shift", "mean", "invstd", "h", "w")


def _forward_check(model, x, train, keep=False):
    """Returns (logits tuple, per-layer records or None).  keep=True retains what the backward pass needs."""
    st = _state(model)
    specs = st.specs
    ops.require_cuda(x, "input")
    if x.dtype == torch.uint8:      # raw NHWC image batch: the device-side input pipeline in its fp32 NCHW form
        x = ops.input_u8(x.contiguous(), nhwc32=False, **(model.__dict__.get("_nunet_input_norm") or {}))
    x = x.float().contiguous()
    N, Cin, H, W = x.shape
    if Cin != specs[0].cin_real or H % 16 or W % 16:
        raise ValueError("expected %d input channels and height/width multiples of 16" % specs[0].cin_real)
    dev = x.device
    f32 = dict(dtype=torch.float32, device=dev)
    P, S = ops._ptr, ops._stream
    xin = torch.empty((N, H, W, Cin), **f32)
    _check_call("nunet_check_nchw_to_nhwc", P(x), N, Cin, H, W, P(xin), S())
    acts, recs = {}, []
    for s in specs:
        h, w = H >> s.level, W >> s.level
        if s.sources[0][0] < 0:
            buf, cin = xin, Cin
        elif s.which == 2:
            buf, cin = acts[s.idx - 1], specs[s.idx - 1].cout
        else:
            cin = s.cin_real
            buf = torch.empty((N, h, w, cin), **f32)
            off = 0
            for pi, kind, c in s.sources:      # torch.cat order: same-level nodes, then the upsampled one (archs1.py:117-132)
                src = acts[pi]
                _check_call("nunet_check_route", P(src), N, src.shape[1], src.shape[2], c, _ROUTE_MODE[kind], P(buf), cin, off, S())
                off += c
        wt = torch.empty(9 * cin * s.cout, **f32)
        y = torch.empty((N, h, w, s.cout), **f32)
        wgt = s.conv.weight.detach().float().contiguous()
        bias = s.conv.bias.detach().float().contiguous() if s.conv.bias is not None else None
        _check_call("nunet_check_conv3x3", P(buf), N, h, w, cin, cin, P(wgt), P(bias), s.cout, P(wt), P(y), S())
        bn = s.bn
        scale, shift = torch.empty(s.cout, **f32), torch.empty(s.cout, **f32)
        mean, invstd = (torch.empty(s.cout, **f32), torch.empty(s.cout, **f32)) if keep else (None, None)
        track = bn.track_running_stats and bn.running_mean is not None
        use_batch = train or not track
        momentum = bn.momentum if bn.momentum is not None else 1.0 / float(int(bn.num_batches_tracked) + 1)
        _check_call("nunet_check_bn", P(y), N * h * w, s.cout, P(bn.weight), P(bn.bias), P(bn.running_mean if (track and (train or not use_batch)) else None),
                    P(bn.running_var if (track and (train or not use_batch)) else None), float(momentum), float(bn.eps), 1 if use_batch else 0,
                    P(scale), P(shift), P(mean), P(invstd), S())
        a = torch.empty((N, h, w, s.cout), **f32)
        _check_call("nunet_check_bn_relu", P(y), N * h * w, s.cout, P(scale), P(shift), P(a), s.cout, 0, S())
        acts[s.idx] = a
        if keep:
            r = _CheckRecord()
            r.buf, r.cin, r.y, r.a, r.scale, r.shift, r.mean, r.invstd, r.h, r.w = buf, cin, y, a, scale, shift, mean, invstd, h, w
            recs.append(r)
    if train:
        nbt = [s.bn.num_batches_tracked for s in specs if s.bn.num_batches_tracked is not None]
        if nbt:
            torch._foreach_add_(nbt, 1)
    outs = []
    for s in specs:
        if s.head is not None:
            K = s.head.out_channels
            a = acts[s.idx]
            logits = torch.empty((N, K, a.shape[1], a.shape[2]), **f32)
            wh = s.head.weight.detach().float().reshape(K, -1).contiguous()
            bh = s.head.bias.detach().float().contiguous() if s.head.bias is not None else None
            _check_call("nunet_check_head", P(a), N, a.shape[1], a.shape[2], s.cout, P(wh), P(bh), K, P(logits), S())
            outs.append(logits)
    return tuple(outs), (recs if keep else None)


def _backward_check(model, N, recs, grad_logits):
    """Autograd of _forward_check in fp32 (fp64 reductions).  grad_logits: {layer idx: dlogits fp32 NCHW}.  Returns {id(param): grad}."""
    specs = _state(model).specs
    dev = recs[0].y.device
    f32 = dict(dtype=torch.float32, device=dev)
    P, S = ops._ptr, ops._stream
    grads, ga = {}, {}
    for s in reversed(specs):
        r = recs[s.idx]
        h, w = r.h, r.w
        g = ga.pop(s.idx, None)
        if g is None:
            g = torch.zeros((N, h, w, s.cout), **f32)
        if s.head is not None and grad_logits.get(s.idx) is not None:
            K = s.head.out_channels
            dl = grad_logits[s.idx].float().contiguous()
            wh = s.head.weight.detach().float().reshape(K, -1).contiguous()
            dwh, dbh = torch.empty((K, s.cout), **f32), torch.empty((K,), **f32)
            _check_call("nunet_check_head_bwd", P(dl), P(r.a), N, h, w, s.cout, P(wh), K, P(g), P(dwh), P(dbh), S())
            grads[id(s.head.weight)] = dwh.view_as(s.head.weight)
            if s.head.bias is not None:
                grads[id(s.head.bias)] = dbh
        dgamma, dbeta = torch.empty(s.cout, **f32), torch.empty(s.cout, **f32)
        gy = torch.empty((N, h, w, s.cout), **f32)
        _check_call("nunet_check_relu_bn_bwd", P(g), P(r.y), N * h * w, s.cout, P(r.scale), P(r.shift), P(s.bn.weight), P(r.mean), P(r.invstd),
                    P(dgamma), P(dbeta), P(gy), S())
        grads[id(s.bn.weight)], grads[id(s.bn.bias)] = dgamma, dbeta
        dw = torch.empty_like(s.conv.weight, dtype=torch.float32)
        db = torch.empty(s.cout, **f32) if s.conv.bias is not None else None
        _check_call("nunet_check_conv3x3_wgrad", P(r.buf), N, h, w, r.cin, r.cin, P(gy), s.cout, P(dw), P(db), S())
        grads[id(s.conv.weight)] = dw
        if db is not None:
            grads[id(s.conv.bias)] = db
        if s.sources[0][0] < 0:
            continue
        wgt = s.conv.weight.detach().float().contiguous()
        wt = torch.empty(9 * r.cin * s.cout, **f32)
        dx = torch.empty((N, h, w, r.cin), **f32)
        _check_call("nunet_check_conv3x3_dgrad", P(gy), N, h, w, s.cout, P(wgt), r.cin, P(wt), P(dx), S())
        off = 0
        for pi, kind, c in s.sources:
            rp = recs[pi]
            if pi not in ga:
                ga[pi] = torch.zeros((N, rp.h, rp.w, c), **f32)
            _check_call("nunet_check_route_bwd", P(dx), N, rp.h, rp.w, c, _ROUTE_MODE[kind], r.cin, off, P(rp.a), P(ga[pi]), S())
            off += c
    return grads


class _CheckFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, *params):
        outs, recs = _forward_check(model, x, True, keep=True)
        ctx.model, ctx.params, ctx.recs, ctx.N = model, params, recs, x.shape[0]
        ctx.head_layers = [s.idx for s in _state(model).specs if s.head is not None]
        return outs

    @staticmethod
    def backward(ctx, *grad_logits):
        gl = {i: g for i, g in zip(ctx.head_layers, grad_logits) if g is not None}
        grads = _backward_check(ctx.model, ctx.N, ctx.recs, gl)
        ctx.recs = None
        return (None, None) + tuple(grads.get(id(p)) if p.requires_grad else None for p in ctx.params)


# ------------------------------------------------------------------------------------------------ a VGGBlock on its own
class _BlockFunction(torch.autograd.Function):
    """conv3x3 -> BN -> ReLU, twice (finished/archs1.py:23-32) as four launches forward (two convs with BatchNorm tails, two activate
    passes) and the matching backward; the block's input is treated as an already activated bf16 NHWC source."""

    @staticmethod
    def forward(ctx, block, x, *params):
        N, Cin, H, W = x.shape
        dev = x.device
        cpad = (Cin + 31) // 32 * 32
        a0 = torch.zeros((N, H, W, cpad), dtype=torch.bfloat16, device=dev)
        a0[..., :Cin] = x.detach().permute(0, 2, 3, 1)          # layout change only: NCHW fp32 -> bf16 NHWC, channels padded to 32
        train = block.training
        st = block.__dict__.setdefault("_nunet_block", {})
        cnt = st.get("counters")
        if cnt is None or cnt.device != dev:
            cnt = st["counters"] = torch.zeros(4, dtype=torch.int32, device=dev)
        recs, src, cin = [], a0, cpad
        for i, (conv, bn) in enumerate(((block.conv1, block.bn1), (block.conv2, block.bn2))):
            cout = conv.out_channels
            if cout % 32:
                raise L.NunetError("a standalone VGGBlock needs output channels in multiples of 32")
            plan = ops.conv_plan(N, H, W, cin, cout)
            wpk = ops.pack_weights(conv.weight, L.PACK_FWD, plan)
            scale, shift = torch.empty(cout, device=dev), torch.empty(cout, device=dev)
            mean = invstd = None
            srcs = [ops.Source(L.SRC_ACT, src, cin)]
            if train:
                mean, invstd = torch.empty(cout, device=dev), torch.empty(cout, device=dev)
                track = bn.track_running_stats and bn.running_mean is not None
                momentum = bn.momentum if bn.momentum is not None else 1.0 / float(int(bn.num_batches_tracked) + 1)
                tail = ops.bn_tail(cnt[i:i + 1], float(N * H * W), bn.weight, bn.bias, conv.bias, bn.running_mean if track else None,
                                   bn.running_var if track else None, momentum, bn.eps, scale, shift, mean, invstd)
                y, _ = ops.conv3x3(N, H, W, srcs, cout, wpk, want_stats=True, tail=tail)
                if bn.num_batches_tracked is not None:
                    bn.num_batches_tracked += 1
            else:
                y, _ = ops.conv3x3(N, H, W, srcs, cout, wpk)
                ops.bn_fold_eval(bn.weight, bn.bias, conv.bias, bn.running_mean, bn.running_var, bn.eps, scale, shift)
            a, _, _, _ = ops.bn_activate(y, scale, shift)
            recs.append((src, cin, y, scale, shift, mean, invstd))
            src, cin = a, cout
        ctx.block, ctx.recs, ctx.shape, ctx.params, ctx.train = block, recs, (N, Cin, H, W), params, train
        return src.permute(0, 3, 1, 2).float()

    @staticmethod
    def backward(ctx, gout):
        block, (N, Cin, H, W) = ctx.block, ctx.shape
        if not ctx.train:
            raise L.NunetError("backward through a standalone VGGBlock needs train mode (batch statistics)")
        dev = gout.device
        cnt = block.__dict__["_nunet_block"]["counters"]
        grads = {}
        g = gout.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)     # gradient w.r.t. the block's output activation
        count = float(N * H * W)
        convs = ((block.conv1, block.bn1), (block.conv2, block.bn2))
        gz2 = None
        for i in (1, 0):
            conv, bn = convs[i]
            src, cin, y, scale, shift, mean, invstd = ctx.recs[i]
            cout = conv.out_channels
            dgam, dbet, k1, k2, k3 = (torch.empty(cout, device=dev) for _ in range(5))
            if i == 1:
                tail = ops.bn_bwd_tail(cnt[2:3], count, bn.weight, mean, invstd, dgam, dbet, k1, k2, k3)
                gz, _ = ops.grad_router(N, H, W, cout, y, scale, shift, mean, invstd, [ops.Consumer(g, cout, 0, L.SRC_SAME)], tail=tail)
            else:
                gz, dgam, dbet, k1, k2, k3 = gz2
            gy = ops.bn_bwd_apply(gz, y, k1, k2, k3, out=gz)
            dw = torch.empty((cout, cin, 3, 3), device=dev)
            ops.conv3x3_wgrad(N, H, W, [ops.Source(L.SRC_ACT, src, cin)], cout, gy, dw)
            grads[id(conv.weight)] = dw[:, :conv.in_channels].contiguous()
            grads[id(bn.weight)], grads[id(bn.bias)] = dgam, dbet
            if conv.bias is not None:
                grads[id(conv.bias)] = torch.zeros_like(conv.bias)     # train-mode BatchNorm cancels the conv bias exactly
            plan = ops.conv_plan(N, H, W, cout, cin)
            wpd = ops.pack_weights(conv.weight, L.PACK_DGRAD, plan)
            if i == 1:      # the first conv's gradient router runs in this dgrad's epilogue
                _, _, y1, sc1, sh1, mu1, rs1 = ctx.recs[0]
                c1 = convs[0][0].out_channels
                d1 = [torch.empty(c1, device=dev) for _ in range(5)]
                bt = ops.bn_bwd_tail(cnt[3:4], count, convs[0][1].weight, mu1, rs1, *d1)
                gz1, _ = ops.conv3x3(N, H, W, [ops.Source(L.SRC_ACT, gy, cout)], cin, wpd, cls="conv_dgrad", mask=ops.Mask(y1, sc1, sh1, mu1, rs1), bwd_tail=bt)
                gz2 = (gz1, *d1)
            else:
                dx, _ = ops.conv3x3(N, H, W, [ops.Source(L.SRC_ACT, gy, cout)], cin, wpd, cls="conv_dgrad")
        gx = dx[..., :Cin].permute(0, 3, 1, 2).float() if ctx.needs_input_grad[1] else None
        ctx.recs = None
        return (None, gx) + tuple(grads.get(id(p)) if p.requires_grad else None for p in ctx.params)


def run_block(block, x):
    """VGGBlock.forward (finished/archs1.py:23-32) outside a UNet / NestedUNet."""
    L.lib()
    ops.require_cuda(x, "input")
    if x.dim() != 4 or x.shape[1] != block.conv1.in_channels:
        raise ValueError("expected NCHW input with %d channels" % block.conv1.in_channels)
    if x.shape[3] + 2 > 256 * 64:
        raise L.NunetError("unsupported width")
    x = x.float().contiguous()
    params = [p for p in block.parameters()]
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)) and block.training:
        return _BlockFunction.apply(block, x, *params)
    with torch.no_grad():
        return _BlockFunction.forward(_NoCtxObj(), block, x, *params)


class _NoCtxObj(object):
    pass


def run_model(model, x):
    """model(input) of the reference (trains.py:117-125, val.py:92-95)."""
    L.lib()  # fail loudly if the CUDA library is missing
    params = [p for p in model.parameters()]
    need_grad = model.training and torch.is_grad_enabled() and any(p.requires_grad for p in params)
    if CHECK_FP32:
        if need_grad:
            outs = _CheckFunction.apply(model, x, *params)
        else:
            with torch.no_grad():
                outs, _ = _forward_check(model, x, model.training)
        return list(outs) if model._returns_list else outs[0]     # check mode: the per-head loss launches (its own reference path)
    if need_grad:
        outs = _NetFunction.apply(model, x, *params)
    elif (GRAPH_EVAL and not model.training and torch.is_tensor(x) and x.is_cuda and x.dim() == 4 and x.dtype in (torch.float32, torch.uint8)
          and x.is_contiguous() and not torch.cuda.is_current_stream_capturing() and _world(model) is None):
        outs = _graphed_eval(model, x, params)
    else:
        with torch.no_grad():
            _, _, logits = _forward(model, x, model.training)
        outs = tuple(t for _, t in logits)
    if model._returns_list:
        import losses
        return losses.mark_head_group(list(outs))     # BCEDiceLoss then evaluates all heads in one launch (losses._group_loss)
    return outs[0]
