"""Drop-in for the reference's metrics.iou_score (metrics.py:6-18).

The reference moves sigmoid(output) and the target to the host and counts with numpy; here the thresholding
(sigmoid_f32(x) > 0.5, target > 0.5) and the two integer counts happen on the device (bit-exact with the reference
given the same logits, including logit == 0 -> False); only the two int64 counts cross to the host, where the ratio is
formed in float64 exactly as metrics.py:18 does."""
import numpy as np
import torch

import ops


def iou_counts_device(output, target):
    """(intersection, union) as an int64[2] tensor ON THE DEVICE — no host synchronisation (utils.StepScalars reads it one step
    late).  The counts BCEDiceLoss cached for this (output, target) pair are reused; otherwise one counting launch."""
    cached = getattr(output, "__dict__", {}).get("_nunet_iou") if torch.is_tensor(output) else None
    if cached is not None and torch.is_tensor(target) and cached[0] == target.data_ptr() and cached[1] == target._version \
            and cached[2] == output._version:
        return cached[3]
    ops.require_cuda(output, "iou_counts_device output")
    o = output.detach().to(dtype=torch.float32).contiguous()
    t = target.detach().to(device=o.device, dtype=torch.float32).contiguous()
    if o.shape != t.shape:
        raise ValueError("output and target must have the same shape")
    return ops.iou_counts(o, t)


def iou_counts(output, target):
    """(intersection, union) as Python ints over the whole batch and all classes (metrics.py:13-16)."""
    cached = getattr(output, "__dict__", {}).get("_nunet_iou") if torch.is_tensor(output) else None
    if cached is not None and torch.is_tensor(target) and cached[0] == target.data_ptr() and cached[1] == target._version \
            and cached[2] == output._version:
        i, u = cached[3].tolist()
        return int(i), int(u)
    if not torch.cuda.is_available():
        raise RuntimeError("metrics.iou_score runs on the B200 path only (no CPU fallback)")
    dev = output.device if (torch.is_tensor(output) and output.is_cuda) else (
        target.device if (torch.is_tensor(target) and target.is_cuda) else torch.device("cuda", torch.cuda.current_device()))
    if torch.is_tensor(output):
        o = output.detach().to(device=dev, dtype=torch.float32).contiguous()
    else:
        # ndarray input is ALREADY sigmoid-ed in the reference's contract (metrics.py:9-10 only applies it to tensors):
        # map probabilities back through the logit of the 0.5 threshold: p > 0.5  <=>  (p - 0.5) > 0
        p = torch.as_tensor(np.asarray(output), dtype=torch.float32, device=dev)
        o = torch.where(p > 0.5, torch.ones_like(p), -torch.ones_like(p)).contiguous()
    t = target.detach() if torch.is_tensor(target) else torch.as_tensor(np.asarray(target))
    t = t.to(device=dev, dtype=torch.float32).contiguous()
    if o.shape != t.shape:
        raise ValueError("output and target must have the same shape")
    i, u = ops.iou_counts(o, t).tolist()
    return int(i), int(u)


def iou_score(output, target):
    smooth = 1e-5
    intersection, union = iou_counts(output, target)
    return (intersection + smooth) / (union + smooth)
