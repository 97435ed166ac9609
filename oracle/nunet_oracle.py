"""fp32 CPU restatement of the reference's UNet++ / U-Net hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function cites the reference lines it restates (paths relative to the reference checkout):
  finished/archs1.py:14-32   VGGBlock      -> vgg_block
  finished/archs1.py:35-71   UNet          -> unet_forward
  finished/archs1.py:74-143  NestedUNet    -> nested_unet_forward
  losses.py:103-117          BCEDiceLoss   -> bce_dice_loss
  losses.py:49-96,120-129    LovaszHingeLoss -> lovasz_hinge_loss
  metrics.py:6-18            iou_score     -> iou_score / iou_counts
  trains.py:113-136          train step    -> train_step
  trains.py:159-177, val.py:87-98  eval    -> eval_step

The arithmetic of conv / batch-norm itself lives in PyTorch (third party, no version pinned by the reference:
requirements.txt lists only albumentations; README says "PyTorch 1.x"); this container's torch 2.11.0 CPU path is
the de-facto oracle.  Convolution is delegated to F.conv2d (fp32, mkldnn); everything else (batch-norm statistics and
running-stat update, ReLU, 2x2 max-pool, align_corners bilinear 2x upsample, concat order, heads, loss, IoU) is
restated explicitly so the CUDA kernels have a line-by-line arithmetic statement to be compared with.

PARITY PINNING: the reference ships no golden vectors or tests for this path.  This oracle is
pinned against the unmodified reference modules imported from /root/reference in the build container
(tests/golden/make_golden.py writes tests/golden/*.pt; tests/test_oracle.py replays them), i.e. against outputs of
the reference itself run here.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

NB_FILTER = (32, 64, 128, 256, 512)  # finished/archs1.py:39,78
BN_EPS = 1e-5                        # nn.BatchNorm2d default, finished/archs1.py:19,21
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------------- graph description
def nested_graph():
    """Node list of NestedUNet in the reference's execution order (finished/archs1.py:114-132).

    Each entry: (node, [(source, kind), ...]) with kind in {'input','same','pool','up'}; concat order is the list order
    (same-level nodes in increasing j, the upsampled lower node LAST)."""
    g = [("0_0", [("input", "input")])]
    order = ["1_0", "0_1", "2_0", "1_1", "0_2", "3_0", "2_1", "1_2", "0_3", "4_0", "3_1", "2_2", "1_3", "0_4"]
    for name in order:
        i, j = int(name[0]), int(name[2])
        if j == 0:
            g.append((name, [("%d_0" % (i - 1), "pool")]))
        else:
            srcs = [("%d_%d" % (i, jj), "same") for jj in range(j)] + [("%d_%d" % (i + 1, j - 1), "up")]
            g.append((name, srcs))
    return g


def unet_graph():
    """UNet (finished/archs1.py:58-68): encoder column plus the last anti-diagonal with 2-source concats."""
    g = [("0_0", [("input", "input")])]
    for i in range(1, 5):
        g.append(("%d_0" % i, [("%d_0" % (i - 1), "pool")]))
    g.append(("3_1", [("3_0", "same"), ("4_0", "up")]))
    g.append(("2_2", [("2_0", "same"), ("3_1", "up")]))
    g.append(("1_3", [("1_0", "same"), ("2_2", "up")]))
    g.append(("0_4", [("0_0", "same"), ("1_3", "up")]))
    return g


# --------------------------------------------------------------------------------------------- primitive ops
def max_pool_2x2(x):
    """nn.MaxPool2d(2, 2) (finished/archs1.py:41,82) on NCHW with even H, W."""
    n, c, h, w = x.shape
    return x.reshape(n, c, h // 2, 2, w // 2, 2).amax(dim=(3, 5))


def upsample_bilinear_2x(x):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) (finished/archs1.py:42,83).

    ATen's area_pixel_compute_scale: scale = (in-1)/(out-1) in fp32; src = scale*dst; i0 = floor(src); i1 = min(i0+1, in-1);
    l1 = src - i0; l0 = 1 - l1; out = l0*x[i0] + l1*x[i1] (rows first, then columns, combined as a 4-tap sum)."""
    n, c, h, w = x.shape

    def taps(size_in):
        size_out = 2 * size_in
        scale = np.float32(size_in - 1) / np.float32(size_out - 1) if size_out > 1 else np.float32(0)
        dst = np.arange(size_out, dtype=np.float32)
        src = scale * dst
        i0 = np.floor(src).astype(np.int64)
        i0 = np.minimum(i0, size_in - 1)
        i1 = np.minimum(i0 + 1, size_in - 1)
        l1 = (src - i0.astype(np.float32)).astype(np.float32)
        l0 = (np.float32(1) - l1).astype(np.float32)
        return torch.from_numpy(i0), torch.from_numpy(i1), torch.from_numpy(l0), torch.from_numpy(l1)

    y0, y1, ly0, ly1 = taps(h)
    x0, x1, lx0, lx1 = taps(w)
    ly0 = ly0.view(1, 1, -1, 1)
    ly1 = ly1.view(1, 1, -1, 1)
    lx0 = lx0.view(1, 1, 1, -1)
    lx1 = lx1.view(1, 1, 1, -1)
    r0 = x[:, :, y0, :]
    r1 = x[:, :, y1, :]
    return ly0 * (lx0 * r0[:, :, :, x0] + lx1 * r0[:, :, :, x1]) + ly1 * (lx0 * r1[:, :, :, x0] + lx1 * r1[:, :, :, x1])


def batch_norm(y, prefix, sd, train, stats_out=None):
    """nn.BatchNorm2d forward (finished/archs1.py:19,21,25,29).

    train: batch mean / biased variance over N*H*W normalise; running stats updated with momentum 0.1 and the
    UNBIASED variance; num_batches_tracked += 1.  eval: running stats normalise."""
    gamma, beta = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if train:
        m = y.shape[0] * y.shape[2] * y.shape[3]
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        if stats_out is not None:
            stats_out[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * sd[prefix + ".running_mean"] + BN_MOMENTUM * mean.detach()
            stats_out[prefix + ".running_var"] = (1 - BN_MOMENTUM) * sd[prefix + ".running_var"] + BN_MOMENTUM * var.detach() * (m / (m - 1))
            stats_out[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
    else:
        mean, var = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    inv = torch.rsqrt(var + BN_EPS)
    return (y - mean.view(1, -1, 1, 1)) * (inv * gamma).view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)


def vgg_block(x, name, sd, train, stats_out=None):
    """VGGBlock.forward (finished/archs1.py:23-32): conv3x3(pad 1, bias) -> BN -> ReLU, twice."""
    p = "conv" + name
    out = F.conv2d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    out = torch.relu(batch_norm(out, p + ".bn1", sd, train, stats_out))
    out = F.conv2d(out, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    out = torch.relu(batch_norm(out, p + ".bn2", sd, train, stats_out))
    return out


def _run_graph(graph, x, sd, train, stats_out):
    feats = {"input": x}
    for node, srcs in graph:
        parts = []
        for s, kind in srcs:
            t = feats[s]
            if kind == "pool":
                t = max_pool_2x2(t)
            elif kind == "up":
                t = upsample_bilinear_2x(t)
            parts.append(t)
        feats[node] = vgg_block(parts[0] if len(parts) == 1 else torch.cat(parts, 1), node, sd, train, stats_out)
    return feats


def nested_unet_forward(sd, x, deep_supervision=False, train=True, stats_out=None):
    """NestedUNet.forward (finished/archs1.py:114-143). `sd` is a reference-compatible state_dict (fp32, OIHW)."""
    f = _run_graph(nested_graph(), x, sd, train, stats_out)
    if deep_supervision:
        return [F.conv2d(f["0_%d" % k], sd["final%d.weight" % k], sd["final%d.bias" % k]) for k in (1, 2, 3, 4)]
    return F.conv2d(f["0_4"], sd["final.weight"], sd["final.bias"])


def unet_forward(sd, x, train=True, stats_out=None):
    """UNet.forward (finished/archs1.py:58-71)."""
    f = _run_graph(unet_graph(), x, sd, train, stats_out)
    return F.conv2d(f["0_4"], sd["final.weight"], sd["final.bias"])


# --------------------------------------------------------------------------------------------- loss / metric
def bce_dice_loss(logits, target):
    """BCEDiceLoss.forward (losses.py:107-117)."""
    x, t = logits, target
    # F.binary_cross_entropy_with_logits, mean reduction: (1-t)*x + softplus(-x)  [= max(x,0) - x*t + log1p(exp(-|x|)),
    # written in the smooth form so that autograd gives sigmoid(x)-t also at x == 0]
    bce = ((1 - t) * x + F.softplus(-x)).mean()
    smooth = 1e-5
    p = torch.sigmoid(x)
    num = t.size(0)
    p = p.reshape(num, -1)
    tt = t.reshape(num, -1)
    inter = (p * tt).sum(1)
    dice = (2.0 * inter + smooth) / (p.sum(1) + tt.sum(1) + smooth)
    dice = 1 - dice.sum() / num
    return 0.5 * bce + dice


def lovasz_hinge_loss(logits, target):
    """LovaszHingeLoss.forward (losses.py:120-129) = lovasz_hinge(input.squeeze(1), target.squeeze(1), per_image=True)
    (losses.py:62-75): per image, errors = 1 - logits*signs sorted descending, the Jaccard gradient of the sorted labels
    (lovasz_grad, losses.py:49-60), loss = dot(relu(errors_sorted), grad) (lovasz_hinge_flat, losses.py:78-96); mean over images."""
    x = logits.squeeze(1)
    t = target.squeeze(1)
    total = 0
    for xi, ti in zip(x, t):
        xi, ti = xi.reshape(-1), ti.reshape(-1)
        signs = 2.0 * ti.float() - 1.0
        errors = 1.0 - xi * signs
        errors_sorted, perm = torch.sort(errors, dim=0, descending=True)
        gt_sorted = ti[perm]
        gts = gt_sorted.sum()
        intersection = gts - gt_sorted.float().cumsum(0)
        union = gts + (1 - gt_sorted).float().cumsum(0)
        jaccard = 1.0 - intersection / union
        if len(gt_sorted) > 1:
            jaccard = torch.cat([jaccard[:1], jaccard[1:] - jaccard[:-1]])
        total = total + torch.dot(torch.relu(errors_sorted), jaccard.detach())
    return total / x.shape[0]


def iou_counts(logits, target):
    """Integer intersection / union counts of metrics.py:13-16 (whole batch, all classes)."""
    o = torch.sigmoid(logits.detach().float()).cpu().numpy() if torch.is_tensor(logits) else np.asarray(logits)
    t = target.detach().cpu().numpy() if torch.is_tensor(target) else np.asarray(target)
    o_ = o > 0.5
    t_ = t > 0.5
    return int((o_ & t_).sum()), int((o_ | t_).sum())


def iou_score(logits, target):
    """metrics.py:6-18."""
    i, u = iou_counts(logits, target)
    smooth = 1e-5
    return (i + smooth) / (u + smooth)


# --------------------------------------------------------------------------------------------- step loop
def loss_and_iou(outputs, target):
    """trains.py:117-127: deep supervision = mean of the per-head losses, IoU on the last head."""
    if isinstance(outputs, (list, tuple)):
        loss = 0
        for o in outputs:
            loss = loss + bce_dice_loss(o, target)
        loss = loss / len(outputs)
        return loss, iou_score(outputs[-1], target)
    return bce_dice_loss(outputs, target), iou_score(outputs, target)


def sgd_step(params, grads, bufs, lr=1e-3, momentum=0.9, weight_decay=1e-4):
    """torch.optim.SGD as configured in trains.py:229-231 (nesterov False, dampening 0)."""
    for k in params:
        g = grads[k] + weight_decay * params[k]
        if bufs.get(k) is None:
            bufs[k] = g.clone()
        else:
            bufs[k].mul_(momentum).add_(g)
        params[k] = params[k] - lr * bufs[k]
    return params, bufs


PARAM_SUFFIXES = (".weight", ".bias")


def train_step(sd, x, target, arch="NestedUNet", deep_supervision=False, lr=1e-3, momentum=0.9, weight_decay=1e-4, bufs=None):
    """One iteration of trains.py:113-136 on a state_dict; returns (new_sd, loss, iou, grads, outputs)."""
    sd = OrderedDict((k, v.clone()) for k, v in sd.items())
    pkeys = [k for k in sd if k.endswith(PARAM_SUFFIXES) and sd[k].is_floating_point() and "running" not in k]
    for k in pkeys:
        sd[k].requires_grad_(True)
    stats = {}
    if arch == "NestedUNet":
        out = nested_unet_forward(sd, x, deep_supervision, True, stats)
    else:
        out = unet_forward(sd, x, True, stats)
    loss, iou = loss_and_iou(out, target)
    grads_list = torch.autograd.grad(loss, [sd[k] for k in pkeys])
    grads = dict(zip(pkeys, grads_list))
    bufs = {} if bufs is None else bufs
    params = {k: sd[k].detach() for k in pkeys}
    params, bufs = sgd_step(params, grads, bufs, lr, momentum, weight_decay)
    new_sd = OrderedDict()
    for k, v in sd.items():
        if k in params:
            new_sd[k] = params[k]
        elif k in stats:
            new_sd[k] = stats[k]
        else:
            new_sd[k] = v.detach()
    outs = [o.detach() for o in out] if isinstance(out, list) else out.detach()
    return new_sd, float(loss.detach()), iou, grads, outs


def eval_step(sd, x, target, arch="NestedUNet", deep_supervision=False):
    """trains.py:159-177 / val.py:87-98 under no_grad with running statistics."""
    with torch.no_grad():
        if arch == "NestedUNet":
            out = nested_unet_forward(sd, x, deep_supervision, False)
        else:
            out = unet_forward(sd, x, False)
        loss, iou = loss_and_iou(out, target)
    return float(loss), iou, out


# --------------------------------------------------------------------------------------------- synthetic data
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def synthetic_batch(n, h, w, num_classes=1, input_channels=3, seed=1, dist="dsb"):
    """Synthetic inputs of the shape/distribution the reference's loader produces.

    'dsb': u8 ~ UniformInt[0,255]; x = ((u8/255 - mean)/std)/255 (albumentations Normalize() trains.py:266, then the
    extra /255 of dataset.py:71).  'normal': x ~ N(0,1).  target = (U[0,1) > 0.5) as float (dataset.py:73)."""
    g = torch.Generator().manual_seed(seed)
    if dist == "dsb":
        u8 = torch.randint(0, 256, (n, input_channels, h, w), generator=g).float()
        mean = torch.tensor([IMAGENET_MEAN[c % 3] for c in range(input_channels)]).view(1, -1, 1, 1)
        std = torch.tensor([IMAGENET_STD[c % 3] for c in range(input_channels)]).view(1, -1, 1, 1)
        x = ((u8 / 255.0 - mean) / std) / 255.0
    else:
        x = torch.randn(n, input_channels, h, w, generator=g)
    t = (torch.rand(n, num_classes, h, w, generator=g) > 0.5).float()
    return x, t
