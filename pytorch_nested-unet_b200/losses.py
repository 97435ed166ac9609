"""Drop-in for the reference's `losses` module on the hot path: BCEDiceLoss (losses.py:103-117).

`criterion(output, target)` returns a 0-dim tensor that supports `+`, `/`, `.backward()`, `.item()` exactly as in
trains.py:120-135.  Forward is ONE fused reduction over (logits, target): BCE-with-logits sum, the per-sample dice sums
and — because metrics.iou_score is always called on the same pair right after (trains.py:124) — the integer
intersection/union counts, which are cached on the logits tensor for metrics.iou_score to pick up.
LovaszHingeLoss (losses.py:49-96,120-129) is a per-image sort + scan kernel (csrc/lovasz.cu)."""
import os
import weakref

import torch
from torch import nn

import _lib as L
import ops

__all__ = ['BCEDiceLoss', 'LovaszHingeLoss']


class _BCEDiceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        loss, sums, counts = ops.loss_iou_fwd(logits, target, want_counts=True)
        ctx.save_for_backward(logits, target, sums)
        ctx.counts = counts
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        logits, target, sums = ctx.saved_tensors
        g = grad_out.contiguous().float()
        return ops.loss_bwd(logits, target, sums, g), None


class _BCEDiceGroupFunction(torch.autograd.Function):
    """The criterion of ALL deep-supervision heads against one target (trains.py:120-123) as one node: one reduction pass forward
    (the target is read once), one launch backward for every head whose loss took part in the total."""

    @staticmethod
    def forward(ctx, target, *logits):
        loss, sums, counts = ops.loss_iou_fwd_multi(list(logits), target)
        ctx.save_for_backward(target, sums, *logits)
        ctx.counts = counts
        return tuple(loss[k] for k in range(len(logits)))

    @staticmethod
    def backward(ctx, *grad_outs):
        target, sums = ctx.saved_tensors[0], ctx.saved_tensors[1]
        logits = list(ctx.saved_tensors[2:])
        gs = [None if g is None else g.contiguous().float() for g in grad_outs]
        return (None,) + tuple(ops.loss_bwd_multi(logits, target, sums, gs))


def mark_head_group(outputs):
    """Called by the engine on the list of deep-supervision logits it returns: BCEDiceLoss evaluates the whole group on the first
    head it is given (one launch for all of them) and hands the other heads their cached value when the loop reaches them.
    The group only holds WEAK references to the logits, and forgets its cached losses once every head has been served: the logits
    must not sit in a reference cycle (tensor -> group -> loss -> autograd node -> saved logits) that only the cyclic garbage
    collector could free — 38 MB of device memory per step at the headline batch."""
    if os.environ.get("NUNET_NO_HEAD_GROUP"):      # A/B switch: the per-head launches
        return outputs
    group = {"tensors": [weakref.ref(o) for o in outputs], "key": None, "losses": None, "served": set()}
    for k, o in enumerate(outputs):
        o.__dict__["_nunet_head_group"] = (group, k)
    return outputs


def _group_loss(input, target, t):
    """The cached / freshly evaluated group loss of `input`, or None when `input` is not a marked head, the group does not qualify
    or has been dissolved (every head served once: a further call takes the ordinary per-head path)."""
    tag = input.__dict__.get("_nunet_head_group")
    if tag is None:
        return None
    group, k = tag
    if group["tensors"] is None:
        return None
    heads = [r() for r in group["tensors"]]
    if any(h is None for h in heads) or not (1 < len(heads) <= ops.MAX_LOSS_HEADS) or heads[k] is not input:
        return None
    key = (target.data_ptr(), target._version, tuple(h._version for h in heads), torch.is_grad_enabled())
    if group["key"] != key:
        ref = heads[0]
        if not all(ops.is_device_tensor(h) and h.dtype == torch.float32 and h.is_contiguous() and h.shape == ref.shape and h.requires_grad == ref.requires_grad
                   for h in heads):
            return None
        if ref.requires_grad and torch.is_grad_enabled():
            losses = _BCEDiceGroupFunction.apply(t, *heads)
            counts = losses[0].grad_fn.counts if losses[0].grad_fn is not None and hasattr(losses[0].grad_fn, "counts") else None
        else:
            lv, _, counts = ops.loss_iou_fwd_multi([h.detach() for h in heads], t)
            losses = tuple(lv[j] for j in range(len(heads)))
        if counts is not None:
            for j, h in enumerate(heads):
                h.__dict__["_nunet_iou"] = (target.data_ptr(), target._version, h._version, counts[j])
        group["key"], group["losses"], group["served"] = key, losses, set()
    out = group["losses"][k]
    group["served"].add(k)
    if len(group["served"]) == len(heads):      # the loop is through: nothing ties the logits to the autograd graph any more
        group["tensors"], group["losses"], group["key"] = None, None, None
    return out


class BCEDiceLoss(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, input, target):
        ops.require_cuda(input, "BCEDiceLoss input")
        if target.shape != input.shape:
            raise ValueError("Target size (%s) must be the same as input size (%s)" % (tuple(target.shape), tuple(input.shape)))
        x = input if (input.dtype == torch.float32 and input.is_contiguous()) else input.float().contiguous()
        t = target.detach()
        t = t if (t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda) else t.to(device=x.device, dtype=torch.float32).contiguous()
        if x is input:
            grouped = _group_loss(input, target, t)
            if grouped is not None:
                return grouped
        if x.requires_grad and torch.is_grad_enabled():
            fn_out = _BCEDiceFunction.apply(x, t)
            counts = fn_out.grad_fn.counts if fn_out.grad_fn is not None and hasattr(fn_out.grad_fn, "counts") else None
        else:
            fn_out, _, counts = ops.loss_iou_fwd(x.detach(), t, want_counts=True)
        if counts is not None:
            # handed to metrics.iou_score(output, target) when it is called with the same tensors
            input.__dict__["_nunet_iou"] = (target.data_ptr(), target._version, input._version, counts)
        return fn_out


class _LovaszFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        loss, coef = ops.lovasz_hinge_fwd(logits, target)
        ctx.save_for_backward(coef)
        ctx.shape = logits.shape
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (coef,) = ctx.saved_tensors
        return ops.lovasz_hinge_bwd(coef, grad_out.contiguous().float(), 1.0 / ctx.shape[0], ctx.shape), None


class LovaszHingeLoss(nn.Module):
    """losses.py:120-129: lovasz_hinge(input.squeeze(1), target.squeeze(1), per_image=True) — per image the Lovasz extension of the
    Jaccard loss over the hinge errors, averaged over the batch."""

    def __init__(self):
        super().__init__()

    def forward(self, input, target):
        ops.require_cuda(input, "LovaszHingeLoss input")
        if target.shape != input.shape:
            raise ValueError("Target size (%s) must be the same as input size (%s)" % (tuple(target.shape), tuple(input.shape)))
        x = input if (input.dtype == torch.float32 and input.is_contiguous()) else input.float().contiguous()
        t = target.detach()
        t = t if (t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda) else t.to(device=x.device, dtype=torch.float32).contiguous()
        if x.requires_grad and torch.is_grad_enabled():
            return _LovaszFunction.apply(x, t)
        return ops.lovasz_hinge_fwd(x.detach(), t)[0]
