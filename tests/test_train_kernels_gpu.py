"""Training-step kernels outside the networks, through the C-ABI: esr_adam_multi (FlatAdam) against torch.optim.Adam, esr_l1_reduce /
esr_l1_grad against nn.L1Loss, esr_bce_rel_loss against the relativistic average BCE written with torch ops
(models/SRRaGAN_model.py:353-354,475-476), and the engines' direct writes into the flat gradient buffer."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(autouse=True)
def _watchdog():
    yield
    from esr_b200 import lib
    wd = lib.watchdog()
    assert wd[0] == 0, 'pipeline watchdog fired: %r' % (wd,)


def _nets():
    torch.manual_seed(0)
    mk = lambda: torch.nn.Sequential(torch.nn.Conv2d(3, 7, 3), torch.nn.Conv2d(7, 5, 3), torch.nn.Linear(9, 3)).to(DEV)
    a, b = mk(), mk()
    b.load_state_dict(a.state_dict())
    return a, b


@pytest.mark.parametrize('wd', [0.0, 1e-2])
def test_flat_adam_matches_torch_adam(wd):
    from esr_b200.optim import FlatAdam, grad_view
    a, b = _nets()
    oa = FlatAdam(a.parameters(), lr=3e-3, betas=(0.9, 0.99), weight_decay=wd).register()
    ob = torch.optim.Adam(b.parameters(), lr=3e-3, betas=(0.9, 0.99), weight_decay=wd)
    assert all(grad_view(p) is not None for p in a.parameters())
    g = torch.Generator(device=DEV).manual_seed(1)
    for it in range(5):
        grads = [torch.randn(p.shape, device=DEV, generator=g) * (10 ** (it - 3)) for p in a.parameters()]
        for net, opt in ((a, oa), (b, ob)):
            opt.zero_grad()
            for p, gr in zip(net.parameters(), grads):
                p.grad = gr.clone()
            opt.step()
        for p, q in zip(a.parameters(), b.parameters()):
            assert torch.allclose(p, q, rtol=2e-6, atol=1e-8), (it, float((p - q).abs().max()))
    # torch.optim.Adam's state-dict layout; a torch Adam resumes from it and vice versa
    sd = oa.state_dict()
    assert set(sd['state'][0].keys()) == {'step', 'exp_avg', 'exp_avg_sq'} and float(sd['state'][0]['step']) == 5
    ob2 = torch.optim.Adam(b.parameters(), lr=3e-3, betas=(0.9, 0.99), weight_decay=wd)
    ob2.load_state_dict(sd)
    oa2 = FlatAdam(a.parameters(), lr=3e-3, betas=(0.9, 0.99), weight_decay=wd)
    oa2.load_state_dict(ob.state_dict())
    grads = [torch.randn(p.shape, device=DEV, generator=g) for p in a.parameters()]
    for net, opt in ((a, oa2), (b, ob2)):
        for p, gr in zip(net.parameters(), grads):
            p.grad = gr.clone()
        opt.step()
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, rtol=2e-6, atol=1e-8)
    # a parameter without a gradient is skipped, like torch does (its moments do not decay)
    for net, opt in ((a, oa2), (b, ob2)):
        opt.zero_grad()
        for k, (p, gr) in enumerate(zip(net.parameters(), grads)):
            p.grad = gr.clone() if k != 2 else None
        opt.step()
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, rtol=2e-6, atol=1e-8)


@pytest.mark.parametrize('shape', [(2, 3, 33, 47), (1, 512, 8, 8), (7,)])
def test_l1_kernels_match_torch(shape):
    from esr_b200.losses import l1_mean
    g = torch.Generator().manual_seed(3)
    a = torch.randn(shape, generator=g).to(DEV).requires_grad_(True)
    b = torch.randn(shape, generator=g).to(DEV)
    b.view(-1)[0] = a.detach().view(-1)[0]            # an exact tie: gradient 0 there, as torch defines it
    ar = a.detach().clone().requires_grad_(True)
    l, lr = l1_mean(a, b), F.l1_loss(ar, b)
    assert abs(float(l) - float(lr)) < 1e-6 * max(1.0, abs(float(lr)))
    (3.0 * l).backward()
    (3.0 * lr).backward()
    assert torch.equal(a.grad, ar.grad)


@pytest.mark.parametrize('n,detach_a', [(4, False), (32, True), (5, False)])
def test_relativistic_bce_matches_torch(n, detach_a):
    from esr_b200.losses import relativistic_bce
    g = torch.Generator().manual_seed(n)
    a0, b0 = (torch.randn(n, 1, generator=g) * 3).to(DEV), (torch.randn(n, 1, generator=g) * 3).to(DEV)
    res = []
    for fused in (True, False):
        a = a0.clone().requires_grad_(not detach_a)
        b = b0.clone().requires_grad_(True)
        if fused:
            la, lb = relativistic_bce(a, b, 1.0, 0.0)
        else:
            la = F.binary_cross_entropy_with_logits(a - b.mean(), torch.ones_like(a))
            lb = F.binary_cross_entropy_with_logits(b - a.mean(), torch.zeros_like(b))
        (0.7 * la + 0.3 * lb).backward()
        res.append((la.detach(), lb.detach(), None if detach_a else a.grad.clone(), b.grad.clone()))
    for x, y in zip(res[0], res[1]):
        if x is not None:
            assert torch.allclose(x, y, rtol=1e-5, atol=1e-7), (x, y)


def test_engine_writes_gradients_into_the_flat_buffer():
    """generator training backward with FlatAdam-registered parameters: the wgrad launches write into the views (.grad IS the view, no
    allocation), a second backward accumulates there, and the values equal the un-registered path bit for bit"""
    import models.modules.architecture as arch
    from CEM.CEMnet import CEMnet, Get_CEM_Conf
    from esr_b200.optim import FlatAdam, grad_view
    torch.manual_seed(2)
    mk = lambda: CEMnet(Get_CEM_Conf(4)).WrapArchitecture_PyTorch(arch.RRDBNet(3, 3, 32, 1, upscale=4, num_latent_channels=0), None).to(DEV).train()
    ma, mb = mk(), mk()
    mb.load_state_dict(ma.state_dict())
    pa = [p for n_, p in ma.named_parameters() if 'Filter_OP' not in n_]
    pb = [p for n_, p in mb.named_parameters() if 'Filter_OP' not in n_]
    opt = FlatAdam(pa, lr=1e-4).register()
    x, hr = torch.rand(2, 3, 24, 40, device=DEV), torch.rand(2, 3, 96, 160, device=DEV)
    with torch.no_grad():
        y0 = ma(x)
    for m in (ma, mb):
        (m(x) - hr).abs().mean().backward()
    for p, q in zip(pa, pb):      # (two launches of the bf16 row kernel differ in summation order: equal to operand precision, not bit for bit)
        assert p.grad is grad_view(p)
        assert float((p.grad - q.grad).norm()) < 3e-2 * float(q.grad.norm())
    flat = opt.flat_grad().clone()
    (ma(x) - hr).abs().mean().backward()          # gradient accumulation: in place, into the same views
    assert float((opt.flat_grad() - 2 * flat).norm()) < 3e-2 * float((2 * flat).norm())      # (bf16 launches: summation order varies)
    assert opt.grads_in_place()
    w0 = pa[0].detach().clone()
    opt.step()
    assert not torch.equal(w0, pa[0].detach()) and float(opt.state[pa[0]]['step']) == 1
    opt.zero_grad()
    assert pa[0].grad is None
    # the engine picks the updated weights up (version counters bumped by the step): same output as a torch.optim.Adam twin
    opt_b = torch.optim.Adam(pb, lr=1e-4)
    for p in pb:
        p.grad = p.grad * 2
    opt_b.step()
    with torch.no_grad():
        y1, y2 = ma(x), mb(x)
    assert torch.isfinite(y1).all() and float((y1 - y2).abs().max()) < 2e-3 and float((y1 - y2).abs().max()) < 0.5 * float((y1 - y0).abs().max())


def test_backward_after_a_second_forward_is_refused():
    """the engine keeps ONE set of saved activations per shape: a stale record raises instead of returning wrong gradients"""
    import models.modules.architecture as arch
    from esr_b200 import lib
    torch.manual_seed(1)
    net = arch.RRDBNet(3, 3, 32, 1, upscale=4, num_latent_channels=0).to(DEV)
    x1 = torch.rand(1, 3, 16, 32, device=DEV).requires_grad_(True)
    x2 = torch.rand(1, 3, 16, 32, device=DEV).requires_grad_(True)
    for p in net.parameters():
        p.requires_grad_(False)
    y1 = net(x1)
    y2 = net(x2)
    y2.sum().backward()
    with pytest.raises(lib.EsrError):
        y1.sum().backward()


def test_deterministic_flag_makes_training_backward_bit_reproducible():
    """esr_set_deterministic(1): one MMA issuer per CTA, no cross-block float atomics - two launches of the same backward agree bit for bit
    (with the default three issuers the fp32 summation order inside an accumulator follows arrival order)"""
    import models.modules.architecture as arch
    from CEM.CEMnet import CEMnet, Get_CEM_Conf
    from esr_b200 import lib
    torch.manual_seed(4)
    model = CEMnet(Get_CEM_Conf(4)).WrapArchitecture_PyTorch(arch.RRDBNet(3, 3, 64, 1, upscale=4, num_latent_channels=0), None).to(DEV).train()
    params = [p for n_, p in model.named_parameters() if 'Filter_OP' not in n_]
    x, hr = torch.rand(2, 3, 40, 136, device=DEV), torch.rand(2, 3, 160, 544, device=DEV)
    lib.check(lib.load().esr_set_deterministic(1))
    try:
        runs = []
        for _ in range(2):
            for p in params:
                p.grad = None
            out = model(x)
            (out - hr).abs().mean().backward()
            runs.append((out.detach().clone(), [p.grad.clone() for p in params]))
        assert torch.equal(runs[0][0], runs[1][0])
        assert all(torch.equal(a, b) for a, b in zip(runs[0][1], runs[1][1]))
    finally:
        lib.check(lib.load().esr_set_deterministic(0))
