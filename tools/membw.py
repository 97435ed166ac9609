"""HBM bandwidth by access mix (torch kernels): read-only (sum), write-only (fill), copy; 1 GiB buffers, best of 5."""
import torch
dev = "cuda:0"
n = 1 << 29
a = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_()
b = torch.empty_like(a)
def best(fn, nbytes):
    t = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return nbytes / min(t) / 1e6
print("read-only  (sum)   %.0f GB/s" % best(lambda: a.view(torch.int16).sum(), n * 2))
print("write-only (fill)  %.0f GB/s" % best(lambda: b.fill_(1.0), n * 2))
print("write-only (memset) %.0f GB/s" % best(lambda: b.zero_(), n * 2))
print("copy (r+w bytes)   %.0f GB/s" % best(lambda: b.copy_(a), n * 4))
