"""Drop-in for the reference's `losses` module on the hot path: BCEDiceLoss (losses.py:103-117).

`criterion(output, target)` returns a 0-dim tensor that supports `+`, `/`, `.backward()`, `.item()` exactly as in
trains.py:120-135.  Forward is ONE fused reduction over (logits, target): BCE-with-logits sum, the per-sample dice sums
and — because metrics.iou_score is always called on the same pair right after (trains.py:124) — the integer
intersection/union counts, which are cached on the logits tensor for metrics.iou_score to pick up.
LovaszHingeLoss (losses.py:49-96,120-129) is a per-image sort + scan kernel (csrc/lovasz.cu)."""
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
