"""WGAN-GP on the B200 (models/modules/loss.py:260-279, models/SRRaGAN_model.py:362-371): the BatchNorm tangent / double-backward
kernels against their torch restatement (tests/disc_emul.py, itself held to the unmodified reference's double backward on the CPU by
tests/test_discriminator_cpu.py), the whole penalty and its parameter gradients against the reference's fixture
(oracle/make_golden_gp.py), and the explorable-SR training configuration (gan_type wgan-gp, non-relativistic) stepping."""
import numpy as np
import pytest
import torch

import disc_emul as E
from util import golden, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(autouse=True)
def _watchdog():
    yield
    from esr_b200 import lib
    wd = lib.watchdog()
    assert wd[0] == 0, 'pipeline watchdog fired: %r' % (wd,)


@pytest.mark.parametrize('n,c,h,w,has_bn', [(4, 16, 8, 8, True), (3, 24, 10, 6, True), (2, 8, 12, 16, False)])
def test_bn_second_order_kernels_match_torch(n, c, h, w, has_bn):
    from esr_b200 import ops
    ops.device_check()
    g = torch.Generator().manual_seed(n * 10 + c)
    y = torch.randn(n, c, h, w, generator=g) * 2 + 0.3
    t = torch.randn(n, c, h, w, generator=g)
    gamma, beta = 1 + 0.3 * torch.randn(c, generator=g), 0.2 * torch.randn(c, generator=g)
    y32, t32 = E.to_planes(y, torch.float32), E.to_planes(t, torch.float32)
    if has_bn:
        mean, invstd, scale, shift = E.bn_stats(y32, c, gamma, beta, 1e-5, 0.1, True, None, None)
    else:
        mean, invstd, scale, shift = torch.zeros(c), torch.ones(c), torch.ones(c), torch.zeros(c)
    d = lambda a: a.to(DEV) if a is not None else None
    for s2d in (False, True):
        r16, rn, rc1, rc2 = E.bn_tangent_fwd(t32, y32, c, scale, shift, mean, invstd, 0.2, torch.float32, has_bn=has_bn, space_to_depth=s2d, want_nchw=True)
        o16, on, c1, c2 = ops.bn_tangent_fwd(d(t32), d(y32), c, d(scale), d(shift), d(mean), d(invstd), 0.2, torch.float16, has_bn=has_bn,
                                             space_to_depth=s2d, want_nchw=True)
        assert rel_err(on.cpu(), rn)[0] < 1e-5 and o16.shape == r16.shape and rel_err(o16.cpu().float(), r16)[0] < 2e-3
        assert (c1.cpu() - rc1).abs().max() < 1e-5 and (c2.cpu() - rc2).abs().max() < 1e-5
    _, _, rc1, rc2 = E.bn_tangent_fwd(t32, y32, c, scale, shift, mean, invstd, 0.2, torch.float32, has_bn=has_bn)
    cpad = E.planes_for(c) * 8
    for layout in (0, 1, 2):
        for have_z in (True, False):
            wb_n, zb_n = torch.randn(n, c, h, w, generator=g), (torch.randn(n, c, h, w, generator=g) if have_z else None)

            def lay(a):
                if a is None:
                    return None
                if layout == 2:
                    return a.contiguous()
                pad = torch.zeros(n, cpad, h, w)
                pad[:, :c] = a
                return E.to_planes(E.s2d_nchw(pad), torch.float32) if layout == 1 else E.to_planes(a, torch.float32)
            dg_r, db_r = torch.zeros(c), torch.zeros(c)
            rt, ry = E.bn_double_bwd(lay(zb_n), lay(wb_n), layout, y32, t32, c, scale, shift, mean, invstd, rc1, rc2, 0.2, torch.float32, has_bn=has_bn,
                                     dgamma=dg_r, dbeta=db_r)
            dg, db = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
            ot, oy = ops.bn_double_bwd(d(lay(zb_n)), d(lay(wb_n)), layout, d(y32), d(t32), c, d(scale), d(shift), d(mean), d(invstd), d(rc1), d(rc2), 0.2,
                                       torch.float16, has_bn=has_bn, dgamma=dg, dbeta=db)
            assert torch.isfinite(ot.float()).all() and torch.isfinite(oy.float()).all() and torch.isfinite(rt).all() and torch.isfinite(ry).all()
            for got, ref, what in ((ot, rt, 'tb'), (oy, ry, 'yb')):
                if float(ref.abs().max()) == 0:          # (no normalisation and no primal adjoint: exactly zero)
                    assert float(got.float().abs().max()) == 0, (what, layout, have_z)
                else:
                    assert rel_err(got.cpu().float(), ref)[0] < 2e-3, (what, layout, have_z, rel_err(got.cpu().float(), ref))
            if has_bn:
                assert rel_err(dg.cpu(), dg_r)[0] < 1e-4
                if have_z:
                    assert rel_err(db.cpu(), db_r)[0] < 1e-4
                else:                       # dbeta = sum of the primal adjoint: exactly zero without one
                    assert float(db.abs().max()) == 0 and float(db_r.abs().max()) == 0


def _critic(fixture_w):
    import models.modules.architecture as arch
    w = golden(fixture_w)
    net = arch.Discriminator_VGG_128(in_nc=3, base_nf=int(w['cfg'][0]), input_patch_size=128)
    sd = {}
    for k in w.files:
        if k.startswith('w:'):
            v = torch.from_numpy(w[k])
            sd[k[2:]] = v.float() if v.dtype.is_floating_point else v
    net.load_state_dict(sd, strict=True)
    return net.to(DEV).train()


@pytest.mark.parametrize('mode,tol', [('parity', 1e-3), ('throughput', 0.3)])
def test_gradient_penalty_matches_reference(mode, tol):
    """penalty value and d(penalty)/d(parameter) for every critic parameter vs the unmodified reference's double backward"""
    from esr_b200 import ops, precision
    from models.modules.loss import GradientPenaltyLoss
    ops.device_check()
    g = golden('wgan_gp_nf8_kf')
    net = _critic('disc_vgg128_nf8_kf')
    interp = torch.from_numpy(g['interp']).to(DEV).requires_grad_(True)
    with precision.use(mode):
        l_gp = GradientPenaltyLoss(device=torch.device(DEV))(interp, net(interp))
        (10.0 * l_gp).backward()
    e_val = abs(float(l_gp) - float(g['l_gp'])) / float(g['l_gp'])
    worst = (0.0, '')
    wscale = max(float(np.abs(g['g:' + n_]).max()) for n_, _ in net.named_parameters())
    for name, p in net.named_parameters():
        ref = 10.0 * torch.from_numpy(g['g:' + name])
        if float(ref.abs().max()) < 1e-6 * 10.0 * wscale:        # exactly-zero gradients (biases the tangent never sees, conv biases in front of BatchNorm)
            assert float(p.grad.abs().max()) < max(tol, 1e-3) * 10.0 * wscale, name
            continue
        worst = max(worst, (rel_err(p.grad.cpu(), ref)[1], name))
    print('gradient penalty [%s]: value %.2e | worst parameter gradient %s rel-L2 %.2e' % (mode, e_val, worst[1], worst[0]))
    assert e_val < tol and worst[0] < tol, (e_val, worst)


def test_wgan_gp_training_steps(tmp_path):
    """options/train/train_explorable_SR.json's GAN settings (gan_type wgan-gp, relativistic 0) through create_model /
    optimize_parameters: the penalty is logged, finite, and both networks move"""
    from esr_b200 import ops
    from models import create_model
    ops.device_check()

    class ND(dict):
        def __missing__(self, k):
            return None
    train = ND(pixel_weight=1e-2, pixel_criterion='l1', gan_type='wgan-gp', gp_weight=10, gan_weight=5e-3, lr_G=1e-4, beta1_G=0.9, weight_decay_G=0,
               lr_D=1e-4, beta1_D=0.9, weight_decay_D=0, D_update_ratio=1, D_init_iters=0, lr_scheme='MultiStepLR', lr_steps=[1000], lr_gamma=0.5,
               grad_accumulation_steps_G=1, grad_accumulation_steps_D=1)
    opt = ND(model='srragan', scale=4, gpu_ids=[0], is_train=True, range=[0, 1], train=train, datasets=ND(train=ND(patch_size=144, batch_size=4)),
             path=ND(models=str(tmp_path / 'models'), pretrained_model_G=None, log=str(tmp_path)),
             network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input=None, latent_input_domain=None, latent_channels=None, norm_type=None,
                          mode='CNA', nf=32, nb=1, in_nc=3, out_nc=3, gc=32, scale=4),
             network_D=ND(which_model_D='discriminator_vgg_128', norm_type='batch', act_type='leakyrelu', mode='CNA', nf=16, in_nc=3, relativistic=0))
    torch.manual_seed(11)
    model = create_model(opt)
    lr = torch.rand(4, 3, 36, 36)
    hr = torch.rand(4, 3, 144, 144)
    d0 = [p.detach().clone() for p in model.netD.parameters()]
    g0 = [p.detach().clone() for p in model.netG.parameters() if p.requires_grad]
    for _ in range(6):
        model.feed_data({'LR': lr, 'HR': hr})
        model.optimize_parameters()
    gp = [v for _, v in model.log_dict['l_d_gp']]
    assert len(gp) >= 5 and all(np.isfinite(v) and v >= 0 for v in gp), gp
    assert any(not torch.equal(a, p.detach()) for a, p in zip(d0, model.netD.parameters()))
    assert any(not torch.equal(a, p.detach()) for a, p in zip(g0, [p for p in model.netG.parameters() if p.requires_grad]))
    assert all(torch.isfinite(p).all() for p in model.netD.parameters())
