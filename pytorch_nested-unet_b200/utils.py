"""Drop-in for the reference's utils module (utils.py:4-33): str2bool, count_params, AverageMeter."""
import argparse


def str2bool(v):
    if v.lower() in ['true', 1]:
        return True
    elif v.lower() in ['false', 0]:
        return False
    else:
        raise argparse.ArgumentTypeError('Boolean value expected.')


def count_params(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


class AverageMeter(object):
    """Batch-size weighted running mean (utils.py:17-33)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class DevicePrefetcher(object):
    """Opt-in wrapper of a loader of (tensor, ..., meta) host batches: the host->device copy of batch i+1 is issued on a copy
    stream before batch i is handed out, so it can overlap batch i's step (the reference's loop copies synchronously,
    trains.py:113-114, and so does train.py here).  Pinned host tensors make the copy asynchronous; pageable ones still work.
    Measured on the B200 box (tools/e2e_ab.py): no faster than the synchronous copy for the 38 MB batches of the headline
    config and noisier, hence not the default."""

    def __init__(self, loader, device=None):
        self.loader, self.device = loader, device

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if self.device is None else torch.device(self.device)
        copy_stream = torch.cuda.Stream(device=dev)
        it = iter(self.loader)

        def load():
            try:
                batch = next(it)
            except StopIteration:
                return None
            with torch.cuda.stream(copy_stream):
                out = tuple(b.to(dev, non_blocking=True) if torch.is_tensor(b) else b for b in batch)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            return out, ev

        nxt = load()
        while nxt is not None:
            out, ev = nxt
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ev)
            for b in out:
                if torch.is_tensor(b):
                    b.record_stream(cur)
            nxt = load()
            yield out


def device_target(target, aug=None):
    """Loss / metric target from what the loader yields: the reference's fp32 NCHW mask tensor is passed through; a raw uint8 NHWC mask
    stack [N,H,W,K] (0 / 255, before dataset.py:73-74's transpose and /255) is converted on the GPU (ops.mask_u8)."""
    import torch
    if torch.is_tensor(target) and target.dtype == torch.uint8:
        import ops
        return ops.mask_u8(target.contiguous(), aug=aug)
    return target
