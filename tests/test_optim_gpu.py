"""The fused multi-tensor optimizers (optim.SGD / optim.Adam, csrc/optim.cu) against torch.optim on identical parameters and
gradients over several steps, including a learning-rate scheduler and state_dict round trip."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SHAPES = [(32, 3, 3, 3), (32,), (64, 32, 3, 3), (1, 32, 1, 1), (1,), (512, 512, 3, 3), (7, 5), (40000,)]


def _params(seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g, device=DEV) * 0.1) for s in SHAPES]


def _grads(params, step, seed=5):
    g = torch.Generator(device=DEV).manual_seed(seed + step)
    for i, p in enumerate(params):
        p.grad = None if (i == 6 and step == 1) else torch.randn(p.shape, generator=g, device=DEV) * 0.01   # one step without a gradient


@pytest.mark.parametrize("kw", [dict(lr=1e-3, momentum=0.9, weight_decay=1e-4), dict(lr=0.05, momentum=0.9, weight_decay=1e-4, nesterov=True),
                                dict(lr=0.01), dict(lr=0.01, momentum=0.5, dampening=0.25, weight_decay=0.0)])
def test_fused_sgd_is_bitwise_torch_sgd(kw):
    import optim
    a, b = _params(1), _params(1)
    oa, ob = optim.SGD(a, **kw), torch.optim.SGD(b, foreach=True, **kw)
    sa = torch.optim.lr_scheduler.CosineAnnealingLR(oa, T_max=3, eta_min=1e-5)
    sb = torch.optim.lr_scheduler.CosineAnnealingLR(ob, T_max=3, eta_min=1e-5)
    for step in range(6):
        _grads(a, step); _grads(b, step)
        oa.step(); ob.step()
        if step % 2 == 1:
            sa.step(); sb.step()
        for x, y in zip(a, b):
            assert torch.equal(x, y), (step, x.shape, float((x - y).abs().max()))
    if kw.get("momentum"):
        for x, y in zip(a, b):
            if "momentum_buffer" in ob.state[y]:
                assert torch.equal(oa.state[x]["momentum_buffer"], ob.state[y]["momentum_buffer"])
    # state_dict round trip into a torch optimizer and back
    import copy
    sd = copy.deepcopy(oa.state_dict())      # state_dict() hands out the live buffers
    c = _params(1)
    oc = optim.SGD(c, **kw)
    oc.load_state_dict(sd)
    for x, y in zip(c, a):
        x.data.copy_(y.data)
    _grads(a, 9); _grads(c, 9)
    oa.step(); oc.step()
    for x, y in zip(a, c):
        assert torch.equal(x, y)


@pytest.mark.parametrize("kw", [dict(lr=1e-3, weight_decay=1e-4), dict(lr=3e-4, betas=(0.8, 0.99), eps=1e-6)])
def test_fused_adam_matches_torch_adam(kw):
    import optim
    a, b = _params(2), _params(2)
    oa, ob = optim.Adam(a, **kw), torch.optim.Adam(b, foreach=True, **kw)
    for step in range(5):
        _grads(a, step); _grads(b, step)
        oa.step(); ob.step()
        for x, y in zip(a, b):
            # same operations in the same order; IEEE division / square root vs the fast-math variants of the foreach kernels: a few ulp
            # (the update is ~lr per element: 1e-8 absolute is 1e-5 of one step's change)
            assert torch.allclose(x, y, rtol=2e-6, atol=1e-8), (step, tuple(x.shape), float((x - y).abs().max()))


def test_no_cpu_fallback():
    import optim
    import _lib as L
    p = [torch.nn.Parameter(torch.zeros(4))]
    p[0].grad = torch.ones(4)
    with pytest.raises(L.NunetError):
        optim.SGD(p, lr=0.1).step()
