"""fp32 torch restatements used by the GPU parity tests (same arithmetic as oracle/nunet_oracle.py, but on the device
and with the bf16 roundings of the CUDA path made explicit so tolerances can be tight)."""
import torch
import torch.nn.functional as F


def bf16r(t):
    return t.to(torch.bfloat16).float()


def nhwc_raw(t_nchw):
    """NCHW fp32 -> bf16 NHWC storage of a raw conv output."""
    return t_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def activated(y_nhwc_bf16, scale, shift):
    """relu(scale*y+shift) in fp32, NCHW."""
    y = y_nhwc_bf16.float().permute(0, 3, 1, 2)
    return torch.relu(y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))


def source_tensor(kind, y_nhwc_bf16, scale, shift):
    a = activated(y_nhwc_bf16, scale, shift)
    if kind == "pool":
        a = F.max_pool2d(a, 2, 2)
    elif kind == "up":
        a = F.interpolate(a, scale_factor=2, mode="bilinear", align_corners=True)
    return a


def conv_ref(x_cat_nchw, w_oihw):
    """bf16-rounded operands, fp32 accumulate (what the tensor core computes), no bias."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return F.conv2d(bf16r(x_cat_nchw), bf16r(w_oihw), None, padding=1)
