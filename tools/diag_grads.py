"""Per-parameter gradient error profile of the CUDA step vs the fp32 CPU oracle (diagnostic, runs on the GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200"))
import archs, losses  # noqa: E402
from oracle import nunet_oracle as O  # noqa: E402

arch, ds, n, h, w = sys.argv[1], sys.argv[2] == "1", int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
torch.manual_seed(0)
model = getattr(archs, arch)(1, 3, ds)
sd = {k: v.clone() for k, v in model.state_dict().items()}
x, t = O.synthetic_batch(n, h, w, 1, 3, 1, "normal")
_, loss_ref, iou_ref, grads_ref, outs_ref = O.train_step(sd, x, t, arch, ds)
model = model.cuda()
crit = losses.BCEDiceLoss()
out = model(x.cuda())
outs = out if isinstance(out, list) else [out]
loss = sum(crit(o, t.cuda()) for o in outs) / len(outs)
loss.backward()
print("loss", float(loss), loss_ref)
refs = outs_ref if isinstance(outs_ref, list) else [outs_ref]
for o, r in zip(outs, refs):
    print("logits relmax %.3e rell2 %.3e" % (float((o.detach().cpu() - r).abs().max() / r.abs().max()), float((o.detach().cpu() - r).norm() / r.norm())))
for k, p in model.named_parameters():
    g = p.grad.cpu()
    r = grads_ref[k]
    if r.norm() < 1e-7:
        continue
    cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
    print("%-24s norm %.4e ref %.4e ratio %.3f rell2 %.3e cos %.4f" % (k, float(g.norm()), float(r.norm()), float(g.norm() / r.norm()), float((g - r).norm() / r.norm()), cos))
