import torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device="cuda").manual_seed(0)
for scale in (1.0, 1e-6):
    for cudnn in (True, False):
        torch.backends.cudnn.enabled = cudnn
        z = torch.randn(4, 32, 96, 96, generator=g, device="cuda", requires_grad=True)
        wh = (torch.randn(3, 32, 1, 1, generator=g, device="cuda") * 0.1).requires_grad_(True)
        dl = torch.randn(4, 3, 96, 96, generator=g, device="cuda") * scale
        a = torch.relu(z)
        o = F.conv2d(a, wh)
        o.backward(dl)
        man = torch.einsum("nkhw,kc->nchw", dl, wh.detach().view(3, 32)) * (z > 0)
        man64 = torch.einsum("nkhw,kc->nchw", dl.double(), wh.detach().view(3, 32).double()) * (z > 0)
        print("scale %g cudnn %s: autograd vs manual rel %.3e ; autograd vs fp64 %.3e ; manual vs fp64 %.3e" % (
            scale, cudnn, float((z.grad - man).norm() / man.norm()), float((z.grad.double() - man64).norm() / man64.norm()),
            float((man.double() - man64).norm() / man64.norm())))
