"""Discriminator_VGG_128 on CPU: the oracle restatement against the reference's golden fixture, the mirror's state-dict
contract, the 4x4-stride-2 -> space-to-depth 3x3 weight re-indexing against torch, and the engine's host logic (with the
CUDA entry points replaced by torch stand-ins on the same layouts) against the reference's outputs and gradients."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import golden, rel_err


def _sd(g):
    out = {}
    for k in g.files:
        if k.startswith('w:'):
            v = torch.from_numpy(g[k])
            out[k[2:]] = v.float() if v.dtype.is_floating_point else v
    return out


def test_oracle_discriminator_matches_reference_golden():
    from oracle import esr_oracle as O
    g = golden('disc_vgg128_nf8')
    sd = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and 'running' not in k else v.clone()) for k, v in _sd(g).items()}
    x = torch.from_numpy(g['x'].astype(np.float32)).requires_grad_(True)
    out = O.discriminator_vgg128_forward(x, sd, training=True, update_running=True)
    assert rel_err(out.detach(), torch.from_numpy(g['out']))[0] < 1e-5
    (out * torch.from_numpy(g['wt'])).sum().backward()
    assert rel_err(x.grad, torch.from_numpy(g['gx']))[0] < 1e-4
    for k in g.files:
        if k.startswith('g:') and not k.endswith('.bias'):
            assert rel_err(sd[k[2:]].grad, torch.from_numpy(g[k]))[0] < 1e-4, k
        if k.startswith('r:'):
            assert rel_err(sd[k[2:]], torch.from_numpy(g[k]))[0] < 1e-5, k
    with torch.no_grad():
        out_eval = O.discriminator_vgg128_forward(x, sd, training=False)
    assert rel_err(out_eval, torch.from_numpy(g['out_eval']))[0] < 1e-5


def test_relativistic_losses_match_torch_definition():
    from oracle import esr_oracle as O
    from models.modules.loss import GANLoss
    torch.manual_seed(0)
    pr, pf = torch.randn(6, 1), torch.randn(6, 1)
    cri = GANLoss('vanilla', 1.0, 0.0)
    d = (cri(pr - pf.mean(), True) + cri(pf - pr.mean(), False)) / 2
    gl = (cri(pr - pf.mean(), False) + cri(pf - pr.mean(), True)) / 2
    assert abs(float(d) - float(O.relativistic_d_loss(pr, pf))) < 1e-6
    assert abs(float(gl) - float(O.relativistic_g_loss(pr, pf))) < 1e-6


def test_mirror_state_dict_matches_reference_keys():
    import models.modules.architecture as arch
    g = golden('disc_vgg128_nf8')
    net = arch.Discriminator_VGG_128(in_nc=3, base_nf=8, input_patch_size=128)
    ref = _sd(g)
    own = net.state_dict()
    assert list(own.keys()) == list(ref.keys())
    for k in own:
        assert tuple(own[k].shape) == tuple(ref[k].shape), k
    net.load_state_dict(ref, strict=True)
    big = arch.Discriminator_VGG_128(in_nc=3, base_nf=64)
    assert sum(p.numel() for p in big.parameters()) == 14502281
    with pytest.raises(NotImplementedError):
        arch.Discriminator_VGG_128(3, 8, num_2_strides=3)
    with pytest.raises(NotImplementedError, match='multiple of 32'):      # 128 - 2 * 20: refused at construction, not at the first step
        arch.Discriminator_VGG_128(3, 8, input_patch_size=88)


def test_k4s2_weight_reindexing_is_exact():
    from esr_b200.disc import k4s2_to_3x3, k3x3_to_k4s2
    from disc_emul import s2d_nchw
    torch.manual_seed(1)
    x = torch.randn(2, 8, 12, 16, dtype=torch.float64)
    w = torch.randn(5, 8, 4, 4, dtype=torch.float64, requires_grad=True)
    ref = F.conv2d(x, w, stride=2, padding=1)
    w3 = k4s2_to_3x3(w.detach()).requires_grad_(True)
    own = F.conv2d(s2d_nchw(x), w3, padding=1)
    assert own.shape == ref.shape and (own - ref).abs().max() < 1e-12
    assert torch.equal(k3x3_to_k4s2(k4s2_to_3x3(w.detach())), w.detach())
    gy = torch.randn_like(ref)
    ref.backward(gy)
    own.backward(gy)
    assert (k3x3_to_k4s2(w3.grad) - w.grad).abs().max() < 1e-12


@pytest.mark.parametrize('training', [True, False])
def test_engine_host_logic_against_reference(monkeypatch, training):
    """DiscEngine with torch stand-ins for the CUDA entry points: logits, running statistics, input and parameter gradients
    against the reference's fixture (train mode) / against the oracle (eval mode)."""
    import disc_emul
    import models.modules.architecture as arch
    from oracle import esr_oracle as O
    disc_emul.install(monkeypatch)
    g = golden('disc_vgg128_nf8')
    net = arch.Discriminator_VGG_128(in_nc=3, base_nf=8, input_patch_size=128)
    net.load_state_dict(_sd(g), strict=True)
    net.compute_dtype = torch.float32
    net.train(training)
    x = torch.from_numpy(g['x'].astype(np.float32)).requires_grad_(True)
    wt = torch.from_numpy(g['wt'])
    out = net(x)
    (out * wt).sum().backward()
    if training:
        assert rel_err(out.detach(), torch.from_numpy(g['out']))[0] < 1e-4
        assert rel_err(x.grad, torch.from_numpy(g['gx']))[0] < 1e-3
        for k in g.files:
            if k.startswith('r:'):
                assert rel_err(net.state_dict()[k[2:]], torch.from_numpy(g[k]))[0] < 1e-5, k
        params = dict(net.named_parameters())
        for k in g.files:
            # conv biases in front of a batch norm have an analytically zero gradient (round-off in the reference too)
            if k.startswith('g:') and not (k.endswith('.bias') and k[2:-5] + '.weight' in params and params[k[2:-5] + '.weight'].dim() == 4
                                            and k != 'g:features.0.bias'):
                assert rel_err(params[k[2:]].grad, torch.from_numpy(g[k]))[0] < 2e-3, k
        assert int(net.features[3].num_batches_tracked) == 1
    else:
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        xr = x.detach().clone().requires_grad_(True)
        ref = O.discriminator_vgg128_forward(xr, sd, training=False)
        (ref * wt).sum().backward()
        assert rel_err(out.detach(), ref.detach())[0] < 1e-4
        assert rel_err(x.grad, xr.grad)[0] < 1e-3


def test_input_gradient_only_when_discriminator_is_frozen(monkeypatch):
    """generator step (SRRaGAN_model.py:421,471): D's parameters do not require grad, the image does"""
    import disc_emul
    import models.modules.architecture as arch
    disc_emul.install(monkeypatch)
    g = golden('disc_vgg128_nf8')
    net = arch.Discriminator_VGG_128(in_nc=3, base_nf=8, input_patch_size=128)
    net.load_state_dict(_sd(g), strict=True)
    net.compute_dtype = torch.float32
    for p in net.parameters():
        p.requires_grad = False
    x = torch.from_numpy(g['x'].astype(np.float32)).requires_grad_(True)
    (net(x) * torch.from_numpy(g['wt'])).sum().backward()
    assert rel_err(x.grad, torch.from_numpy(g['gx']))[0] < 1e-3
    assert all(p.grad is None for p in net.parameters())
    with torch.no_grad():
        assert net(x.detach()).shape == (4, 1)


def test_generic_double_backward_fails_loudly(monkeypatch):
    """a generic second differentiation THROUGH the engine's backward is an error, never a silently missing gradient (the supported
    second-order use, WGAN-GP, goes through GradientPenaltyLoss -> esr_b200.disc.gradient_penalty)"""
    import disc_emul
    import models.modules.architecture as arch
    disc_emul.install(monkeypatch)
    g = golden('disc_vgg128_nf8')
    net = arch.Discriminator_VGG_128(in_nc=3, base_nf=8, input_patch_size=128)
    net.load_state_dict(_sd(g), strict=True)
    net.compute_dtype = torch.float32
    x = torch.from_numpy(g['x'].astype(np.float32)).requires_grad_(True)
    gx = torch.autograd.grad(net(x).sum(), x, create_graph=True)[0]
    with pytest.raises(RuntimeError):
        ((gx.flatten(1).norm(2, dim=1) - 1) ** 2).mean().backward()
    assert net.supports_double_backward is True


def test_gradient_penalty_host_logic_against_reference(monkeypatch):
    """GradientPenaltyLoss on the engine (tangent forward + backward over the primal / tangent pair, esr_b200.disc) against the UNMODIFIED
    reference's double backward (oracle/make_golden_gp.py): penalty value and the gradient of every critic parameter.  The CUDA
    kernels are replaced by torch stand-ins here (tests/disc_emul.py); the GPU suite runs the same comparison on the kernels."""
    import disc_emul
    import models.modules.architecture as arch
    from models.modules.loss import GradientPenaltyLoss
    disc_emul.install(monkeypatch)
    g = golden('wgan_gp_nf8_kf')
    w = golden('disc_vgg128_nf8_kf')
    net = arch.Discriminator_VGG_128(in_nc=3, base_nf=8, input_patch_size=128)
    net.load_state_dict(_sd(w), strict=True)
    net.compute_dtype = torch.float32
    net.train()
    interp = torch.from_numpy(g['interp']).requires_grad_(True)
    l_gp = GradientPenaltyLoss()(interp, net(interp))
    assert abs(float(l_gp) - float(g['l_gp'])) < 1e-4 * float(g['l_gp'])
    (10.0 * l_gp).backward()
    worst = 0.0
    for name, p in net.named_parameters():
        ref = 10.0 * torch.from_numpy(g['g:' + name])
        scale = float(ref.abs().max())
        if scale < 1e-9:
            assert float(p.grad.abs().max()) < 1e-7, name
            continue
        worst = max(worst, rel_err(p.grad, ref)[0])
        assert rel_err(p.grad, ref)[0] < 2e-3, (name, rel_err(p.grad, ref))
    print('gradient penalty, host logic vs reference: worst parameter gradient error %.2e' % worst)


def test_wgan_gp_model_constructs(tmp_path):
    """the reference's explorable-SR training configuration (gan_type wgan-gp, options/train/train_explorable_SR.json:87) constructs"""
    class ND(dict):
        def __missing__(self, k):
            return None
    from models import create_model
    train = ND(pixel_weight=1.0, pixel_criterion='l1', gan_type='wgan-gp', gp_weight=10, gan_weight=5e-3, lr_G=1e-4, beta1_G=0.9, lr_D=1e-4, beta1_D=0.9,
               lr_scheme='MultiStepLR', lr_steps=[10], lr_gamma=0.5, grad_accumulation_steps_G=1, grad_accumulation_steps_D=1)
    opt = ND(model='srragan', scale=4, gpu_ids=None, is_train=True, range=[0, 1], train=train, datasets=ND(train=ND(patch_size=144, batch_size=2)),
             path=ND(models=str(tmp_path / 'models'), pretrained_model_G=None, log=str(tmp_path)),
             network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input=None, latent_input_domain=None, latent_channels=None, norm_type=None,
                          mode='CNA', nf=8, nb=1, in_nc=3, out_nc=3, gc=32, scale=4),
             network_D=ND(which_model_D='discriminator_vgg_128', norm_type='batch', act_type='leakyrelu', mode='CNA', nf=8, in_nc=3, relativistic=0))
    if not torch.cuda.is_available():
        model = create_model(opt)
        assert model.cri_gp is not None and model.l_gp_w == 10


def test_engine_follows_weight_updates_in_place(monkeypatch):
    """after an optimizer step the critic's packed weight objects are refreshed in place (same objects, new values): the
    second forward must see the new weights"""
    import disc_emul
    import models.modules.architecture as arch
    from oracle import esr_oracle as O
    disc_emul.install(monkeypatch)
    g = golden('disc_vgg128_nf8')
    net = arch.Discriminator_VGG_128(in_nc=3, base_nf=8, input_patch_size=128)
    net.load_state_dict(_sd(g), strict=True)
    net.compute_dtype = torch.float32
    net.train()
    x = torch.from_numpy(g['x'].astype(np.float32))
    with torch.no_grad():
        out0 = net(x)
        pk0 = list(net.engine()._pk)
        for p in net.parameters():
            if p.dim() == 4:
                p.mul_(1.05)
        out1 = net(x)
        assert all(a is b for a, b in zip(pk0, net.engine()._pk))           # same packed objects
        ref1 = O.discriminator_vgg128_forward(x, {k: v.detach().clone() for k, v in net.state_dict().items()}, training=True)
    assert rel_err(out1, ref1)[0] < 1e-4
    assert not torch.allclose(out0, out1)
