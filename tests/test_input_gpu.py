"""Device-side input pipeline (csrc/input.cu) against the host expression of the reference's loader: albumentations Normalize()
(trains.py:266) then dataset.py:71's /255, mask/255 (dataset.py:73), np.rot90 + flips (trains.py:257-258)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def _host_norm(u8_nhwc):
    x = u8_nhwc.permute(0, 3, 1, 2).float()
    c = x.shape[1]
    mean = torch.tensor([MEAN[i % 3] for i in range(c)]).view(1, c, 1, 1)
    std = torch.tensor([STD[i % 3] for i in range(c)]).view(1, c, 1, 1)
    return ((x / 255.0 - mean) / std) / 255.0        # oracle.synthetic_batch's expression


@pytest.mark.parametrize("N,H,W,C", [(3, 32, 32, 3), (2, 48, 80, 3), (2, 16, 16, 1), (1, 32, 32, 5)])
def test_u8_normalisation_is_bit_exact_in_fp32_and_rounded_once_to_bf16(N, H, W, C):
    import ops
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (N, H, W, C), generator=g, dtype=torch.uint8)
    ref = _host_norm(u8)
    f32 = ops.input_u8(u8.to(DEV), nhwc32=False)
    assert torch.equal(f32.cpu(), ref)                                      # IEEE divisions in the loader's order
    packed = ops.input_u8(u8.to(DEV), nhwc32=True)
    assert packed.shape == (N, H, W, 32)
    assert torch.equal(packed[..., :C].cpu(), ref.permute(0, 2, 3, 1).to(torch.bfloat16))
    assert float(packed[..., C:].abs().max()) == 0.0
    # the repack of the reference's own fp32 NCHW tensor gives the same operand
    assert torch.equal(packed, ops.input_to_nhwc32(ref.contiguous().to(DEV)))


def test_mask_and_geometric_augmentation_match_numpy_semantics():
    import ops
    g = torch.Generator().manual_seed(4)
    N, S = 16, 32
    img = torch.randint(0, 256, (N, S, S, 3), generator=g, dtype=torch.uint8)
    msk = (torch.rand(N, S, S, 2, generator=g) > 0.5).to(torch.uint8) * 255
    aug = torch.arange(N, dtype=torch.uint8)          # every combination of k (bits 0-1), horizontal (bit 2) and vertical (bit 3) flip
    got_x = ops.input_u8(img.to(DEV), aug=aug.to(DEV), nhwc32=False).cpu()
    got_m = ops.mask_u8(msk.to(DEV), aug=aug.to(DEV)).cpu()
    for n in range(N):
        a = int(aug[n])
        def tf(t):                                    # albumentations: RandomRotate90 (np.rot90) then Flip
            t = torch.rot90(t, a & 3, dims=(0, 1))
            if a & 4:
                t = torch.flip(t, dims=(1,))
            if a & 8:
                t = torch.flip(t, dims=(0,))
            return t
        assert torch.equal(got_x[n], _host_norm(tf(img[n])[None])[0]), n
        assert torch.equal(got_m[n], tf(msk[n]).permute(2, 0, 1).float() / 255.0), n
    plain = ops.mask_u8(msk.to(DEV)).cpu()
    assert torch.equal(plain, msk.permute(0, 3, 1, 2).float() / 255.0)


def test_model_takes_raw_uint8_batches():
    """model(uint8 NHWC) == model(the loader's fp32 NCHW tensor) bit for bit (train and eval mode), gradients included."""
    import archs
    import losses
    g = torch.Generator().manual_seed(9)
    u8 = torch.randint(0, 256, (4, 32, 32, 3), generator=g, dtype=torch.uint8)
    t = (torch.rand(4, 1, 32, 32, generator=g) > 0.5).float().to(DEV)
    crit = losses.BCEDiceLoss()
    res = []
    for inp in (u8.to(DEV), _host_norm(u8).to(DEV)):
        torch.manual_seed(0)
        m = archs.NestedUNet(1, 3, True).to(DEV).train()
        out = m(inp)
        loss = sum(crit(o, t) for o in out) / 4
        loss.backward()
        m.eval()
        with torch.no_grad():
            oe = m(inp)
        res.append(([o.detach().clone() for o in out], [p.grad.clone() for p in m.parameters()], oe[-1].clone()))
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b)
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.equal(a, b)
    assert torch.equal(res[0][2], res[1][2])
