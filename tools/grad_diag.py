"""Per-parameter gradient deviation of one golden case (norm and cosine against the reference's fp32 gradients)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import goldens
import test_model_gpu as T
name = sys.argv[1] if len(sys.argv) > 1 else "nested_wods_n2_32_dsb"
g = goldens.case(name)
model, opt, crit = T._build(g)
x, t = T._synthetic(g)
x, t = x.to(T.DEV), t.to(T.DEV)
model.train()
out = model(x)
loss = sum(crit(o, t) for o in out) / len(out) if isinstance(out, list) else crit(out, t)
loss.backward()
named = dict(model.named_parameters())
rows = []
for k, n_ref in g["grad_norms"].items():
    if k.endswith(("conv1.bias", "conv2.bias")):
        continue
    gr = named[k].grad
    rel = abs(float(gr.norm()) - n_ref) / max(n_ref, 1e-12)
    cos = None
    if k in g["grads"]:
        gk, gref = gr.cpu(), g["grads"][k]
        cos = float((gk * gref).sum() / (gk.norm() * gref.norm() + 1e-30))
    rows.append((rel, k, cos))
rows.sort(reverse=True)
for rel, k, cos in rows[:12]:
    print("%.4f %s cos=%s" % (rel, k, cos))
print("median", sorted(r[0] for r in rows)[len(rows) // 2])
