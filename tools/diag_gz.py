"""Per-layer comparison of g_z (post-ReLU-mask gradient) and g_y (post-BN-backward) between the engine and torch
autograd on the forced forward."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch_nested-unet_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import archs, engine, losses, ops, _lib as L
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
K, n, h, w, ds = 3, 4, 96, 96, True
torch.manual_seed(0)
model = archs.NestedUNet(K, 3, ds).cuda()
gen = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(n, 3, h, w, generator=gen, device="cuda")
t = (torch.rand(n, K, h, w, generator=gen, device="cuda") > 0.5).float()
crit = losses.BCEDiceLoss()
cap = {}
orig_router, orig_bnb = ops.grad_router, ops.bn_backward_reduce
order = []
def router(N, H, W, Cc, y, *a, **kw):
    gz, part = orig_router(N, H, W, Cc, y, *a, **kw)
    order.append((y.data_ptr(), gz))
    if kw.get("dlogits") is not None and "dl" not in cap:
        cap["dl"], cap["wh"], cap["y"], cap["scale"], cap["shift"], cap["gz"] = kw["dlogits"].clone(), kw["wh"].clone(), y, a[0], a[1], gz
    return gz, part
def bnb(part, count, gamma, mean, invstd, ws, dgamma, dbeta, k1, k2, k3, **kw):
    orig_bnb(part, count, gamma, mean, invstd, ws, dgamma, dbeta, k1, k2, k3, **kw)
    order[-1] = order[-1] + (k1, k2, k3)
engine.ops.grad_router = router
engine.ops.bn_backward_reduce = bnb
out = model(x)
loss = sum(crit(o, t) for o in out) / len(out)
loss.backward()
model.zero_grad()
with torch.no_grad():
    _, recs, _ = engine._forward(model, x, True)
by_ptr = {}
specs = engine._state(model).specs
# second forward allocates new y buffers; match by layer order instead (reverse order of specs)
eng = {}
for s, rec in zip(reversed(specs), order):
    eng[s.idx] = rec
def ste(tn): return tn + (tn.to(torch.bfloat16).float() - tn).detach()
acts, ys, zs = {}, {}, {}
for s in specs:
    parts = []
    for pi, kind, c in s.sources:
        if pi < 0: parts.append(x); continue
        a = acts[pi]
        if kind == L.SRC_POOL: a = F.max_pool2d(a, 2, 2)
        elif kind == L.SRC_UP: a = F.interpolate(a, scale_factor=2, mode="bilinear", align_corners=True)
        parts.append(a)
    X = parts[0] if len(parts) == 1 else torch.cat(parts, 1)
    y = F.conv2d(ste(X), ste(s.conv.weight), s.conv.bias, padding=1)
    y_eng = recs[s.idx].y.float().permute(0, 3, 1, 2) + s.conv.bias.detach().view(1, -1, 1, 1)
    y = y + (y_eng - y).detach(); y.retain_grad(); ys[s.idx] = y
    z = F.batch_norm(y, None, None, s.bn.weight, s.bn.bias, True, 0.0, s.bn.eps); z.retain_grad(); zs[s.idx] = z
    acts[s.idx] = torch.relu(z)
routs = [F.conv2d(acts[s.idx], s.head.weight, s.head.bias) for s in specs if s.head is not None]
for r_ in routs: r_.retain_grad()
def ref_loss(o):
    bce = F.binary_cross_entropy_with_logits(o, t)
    p = torch.sigmoid(o).view(n, -1); tt = t.view(n, -1)
    dice = (2 * (p * tt).sum(1) + 1e-5) / (p.sum(1) + tt.sum(1) + 1e-5)
    return 0.5 * bce + 1 - dice.sum() / n
(sum(ref_loss(o) for o in routs) / len(routs)).backward()
for s in reversed(specs):
    _, gz, k1, k2, k3 = eng[s.idx]
    gz_e = gz.float().permute(0, 3, 1, 2)
    gz_t = zs[s.idx].grad
    yb = recs[s.idx].y.float().permute(0, 3, 1, 2)
    gy_e = gz_e * k1.view(1, -1, 1, 1) + yb * k2.view(1, -1, 1, 1) + k3.view(1, -1, 1, 1)
    gy_t = ys[s.idx].grad
    print("%-4s w%d  gz rel %.3e   gy rel %.3e   |gz| %.2e" % (s.node, s.which, float((gz_e - gz_t).norm() / gz_t.norm()),
          float((gy_e - gy_t).norm() / gy_t.norm()), float(gz_t.abs().mean())))

# ---- head path drill-down on the last layer (0_4 conv2)
print("---- drill down")
s = specs[-1]
rec = None

dl_e, wh = cap["dl"], cap["wh"]
dl_t = routs[-1].grad
print("dlogits engine vs torch rel %.3e  (|dl| %.2e)" % (float((dl_e - dl_t).norm() / dl_t.norm()), float(dl_t.abs().mean())))
yv = cap["y"].float().permute(0, 3, 1, 2)
z = yv * cap["scale"].view(1, -1, 1, 1) + cap["shift"].view(1, -1, 1, 1)
ga = torch.einsum("nkhw,kc->nchw", dl_e, wh)
gz_manual = ga * (z > 0)
gz_e = cap["gz"].float().permute(0, 3, 1, 2)
print("router gz vs manual(dl_e*wh*mask) rel %.3e" % float((gz_e - gz_manual).norm() / gz_manual.norm()))
gz_t = zs[specs[-1].idx].grad
print("manual vs torch gz rel %.3e" % float((gz_manual - gz_t).norm() / gz_t.norm()))
ga_t = torch.einsum("nkhw,kc->nchw", dl_t, wh) * (z > 0)
print("torch-dl manual vs torch gz rel %.3e" % float((ga_t - gz_t).norm() / gz_t.norm()))
diff = (gz_t - ga_t)
zt = zs[specs[-1].idx].detach()
print("mask mismatches torch-z vs engine-z:", int(((zt > 0) != (z > 0)).sum()), "of", z.numel())
print("max|z_t - z_e| %.3e" % float((zt - z).abs().max()))
nz = diff.abs() > 1e-3 * ga_t.abs().max()
print("pixels with significant diff:", int(nz.sum()), "; of which engine-mask on:", int((nz & (z > 0)).sum()))
per_c = ((diff ** 2).sum((0, 2, 3)).sqrt() / ((gz_t ** 2).sum((0, 2, 3)).sqrt() + 1e-30))
print("per-channel rel diff:", [round(float(v), 3) for v in per_c])
ratio = (gz_t / ga_t)[(z > 0) & (ga_t.abs() > 1e-9)]
print("ratio gz_t/ga_t quantiles:", [float(torch.quantile(ratio[:100000], q)) for q in (0.01, 0.25, 0.5, 0.75, 0.99)])
