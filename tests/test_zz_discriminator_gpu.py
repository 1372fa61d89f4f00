"""Discriminator_VGG_128 on the B200, through the C-ABI: the BatchNorm / space-to-depth / Linear kernels against torch stand-ins
(tests/disc_emul.py) on identical operands, the whole critic against the unmodified reference's golden fixture
(oracle/make_golden_disc.py) and against the oracle at the reference's size (base_nf 64), and the GAN branch of
SRRaGANModel.optimize_parameters.  Tolerances are stated at the tests."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

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


def _sd(g):
    out = {}
    for k in g.files:
        if k.startswith('w:'):
            v = torch.from_numpy(g[k])
            out[k[2:]] = v.float() if v.dtype.is_floating_point else v
    return out


@pytest.mark.parametrize('n,c,h,w,train', [(4, 16, 8, 8, True), (3, 24, 10, 6, True), (2, 64, 64, 64, True), (2, 8, 4, 4, False)])
def test_bn_kernels_match_torch(n, c, h, w, train):
    from esr_b200 import ops
    ops.device_check()
    g = torch.Generator().manual_seed(n * 100 + c)
    y = torch.randn(n, c, h, w, generator=g) * 2 + 0.5
    gamma, beta = 1 + 0.3 * torch.randn(c, generator=g), 0.2 * torch.randn(c, generator=g)
    rm, rv = 0.1 * torch.randn(c, generator=g), 1 + 0.2 * torch.rand(c, generator=g)
    y32 = E.to_planes(y, torch.float32)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    ref = E.bn_stats(y32, c, gamma, beta, 1e-5, 0.1, train, rm_ref, rv_ref)
    rm_d, rv_d = rm.to(DEV), rv.to(DEV)
    own = ops.bn_stats(y32.to(DEV), c, gamma.to(DEV), beta.to(DEV), 1e-5, 0.1, train, rm_d, rv_d)
    for a, b in zip(own, ref):
        assert rel_err(a.cpu(), b)[0] < 1e-5
    assert rel_err(rm_d.cpu(), rm_ref)[0] < 1e-5 and rel_err(rv_d.cpu(), rv_ref)[0] < 1e-5
    mean, invstd, scale, shift = ref
    dmean, dinv, dscale, dshift = [t.to(DEV) for t in ref]
    # forward apply: plain, space-to-depth, NCHW
    for s2d in (False, True):
        r16, rn = E.bn_lrelu_fwd(y32, c, scale, shift, 0.2, torch.float16, space_to_depth=s2d, want_nchw=True)
        o16, on = ops.bn_lrelu_fwd(y32.to(DEV), c, dscale, dshift, 0.2, torch.float16, space_to_depth=s2d, want_nchw=True)
        assert o16.shape == r16.shape and rel_err(o16.cpu().float(), r16.float())[0] < 1e-3
        assert rel_err(on.cpu(), rn)[0] < 1e-6
    # backward: gradient in the three layouts
    gg = torch.randn(n, c, h, w, generator=g)
    cpad = E.planes_for(c) * 8
    gpad = torch.zeros(n, cpad, h, w)
    gpad[:, :c] = gg
    for layout, gt in ((0, E.to_planes(gg, torch.float32)), (1, E.to_planes(E.s2d_nchw(gpad), torch.float32)), (2, gg.contiguous())):
        dg_r, db_r = torch.zeros(c), torch.zeros(c)
        ref_gy = E.bn_lrelu_bwd(gt, layout, y32, c, scale, shift, mean, invstd, 0.2, torch.float32, has_bn=True, train=train, dgamma=dg_r, dbeta=db_r)
        dg, db = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
        own_gy = ops.bn_lrelu_bwd(gt.to(DEV), layout, y32.to(DEV), c, dscale, dshift, dmean, dinv, 0.2, torch.bfloat16, has_bn=True, train=train,
                                  dgamma=dg, dbeta=db)
        assert rel_err(own_gy.cpu().float(), ref_gy)[0] < 5e-3, layout      # bf16 store
        assert rel_err(dg.cpu(), dg_r)[0] < 1e-4 and rel_err(db.cpu(), db_r)[0] < 1e-4, layout
    # activation-only layer (conv0)
    one, zero = torch.ones(c), torch.zeros(c)
    ref_gy = E.bn_lrelu_bwd(E.to_planes(gg, torch.float32), 0, y32, c, one, zero, zero, one, 0.2, torch.float32, has_bn=False)
    own_gy = ops.bn_lrelu_bwd(E.to_planes(gg, torch.float32).to(DEV), 0, y32.to(DEV), c, one.to(DEV), zero.to(DEV), zero.to(DEV), one.to(DEV), 0.2,
                              torch.float16, has_bn=False)
    assert rel_err(own_gy.cpu().float(), ref_gy)[0] < 1e-3


def test_space_to_depth_is_a_permutation():
    from esr_b200 import ops
    ops.device_check()
    x = torch.randn(2, 16, 6, 10).half()
    own = ops.space_to_depth(E.to_planes(x, torch.float16).to(DEV)).cpu()
    assert torch.equal(own, E.to_planes(E.s2d_nchw(x), torch.float16))


@pytest.mark.parametrize('b,k,j', [(4, 8192, 100), (3, 100, 1), (11, 257, 5)])
def test_linear_kernels_match_torch(b, k, j):
    from esr_b200 import ops
    ops.device_check()
    g = torch.Generator().manual_seed(b + k + j)
    x, w, bias = torch.randn(b, k, generator=g), torch.randn(j, k, generator=g) / k ** 0.5, torch.randn(j, generator=g)
    out = ops.linear_fwd(x.to(DEV), w.to(DEV), bias.to(DEV), lrelu=True)
    ref = F.leaky_relu(F.linear(x.double(), w.double(), bias.double()), 0.2)
    assert rel_err(out.cpu(), ref)[0] < 1e-5
    gy = torch.randn(b, j, generator=g)
    gx, dw, db = ops.linear_bwd(gy.to(DEV), out, x.to(DEV), w.to(DEV))
    rgx, rdw, rdb = E.linear_bwd(gy.double(), ref, x.double(), w.double())
    assert rel_err(gx.cpu(), rgx)[0] < 1e-5 and rel_err(dw.cpu(), rdw)[0] < 1e-5 and rel_err(db.cpu(), rdb)[0] < 1e-5
    gx2, dw2, _ = ops.linear_bwd(gy.to(DEV), None, x.to(DEV), w.to(DEV), want_w=False)
    assert dw2 is None and rel_err(gx2.cpu(), gy.double() @ w.double())[0] < 1e-5


def _mirror(g, dtype):
    import models.modules.architecture as arch
    nf = int(g['cfg'][0])
    net = arch.Discriminator_VGG_128(in_nc=3, base_nf=nf, input_patch_size=128)
    net.load_state_dict(_sd(g), strict=True)
    net.compute_dtype = dtype
    return net.to(DEV)


def _run(net, g, gscale):
    """logits, image gradient and parameter gradients of loss = sum(logits * wt) on the fixture's batch"""
    dev = next(net.parameters()).device
    x = torch.from_numpy(g['x'].astype(np.float32)).to(dev).requires_grad_(True)
    out = net(x)
    (out * (gscale * torch.from_numpy(g['wt']).to(dev))).sum().backward()
    grads = {name: p.grad.detach().cpu() / gscale for name, p in net.named_parameters()}
    return out.detach().cpu(), x.grad.detach().cpu() / gscale, grads, {k: v.detach().cpu().clone() for k, v in net.state_dict().items() if 'running' in k}


# vs the unmodified reference (fp32): logits, gradient rel-L2, gradient cosine; vs the same arithmetic with operands rounded on the host
@pytest.mark.parametrize('dtype,tol_out,tol_l2,min_cos,tol_same', [(torch.float16, 1e-2, 8e-2, 0.995, 5e-2), (torch.bfloat16, 5e-2, 0.3, 0.95, 0.15)])
def test_discriminator_matches_reference_golden(monkeypatch, dtype, tol_out, tol_l2, min_cos, tol_same):
    """Two references.  (1) The unmodified reference's fp32 outputs / gradients (golden fixture).  This random-initialised,
    batch-normalised 10-stage critic on a 4-image batch amplifies a relative perturbation about 100x (fp32 stand-ins reproduce
    the fixture to 1e-6; rounding only the conv operands to fp16 / bf16 on the host moves the gradients by 5e-2 / 1.8e-1 -
    tools/disc_parity.py), so the fp32 comparison is held to operand-precision bounds.  (2) The same algebra with
    the same operand rounding done by torch on the host (tests/disc_emul.py).  Layer 0 agrees to 4e-7; from there on 1e-7-size
    accumulation-order differences flip rounding ties (6e-4 of the fp16 activations after layer 0, 70 % after layer 9 -
    tools/disc_debug.py), each flip a full-ulp perturbation, so deep in the chain the two are independent realisations of the
    same rounding noise: the agreement is held to about half the fp32 bound.  Kernel-level exactness on identical operands is
    what the kernel tests above and tests/test_wgrad_gpu.py / test_conv_rows_gpu.py pin (1e-5)."""
    from esr_b200 import ops
    import models.modules.architecture as arch
    ops.device_check()
    g = golden('disc_vgg128_nf8')
    gscale = 64.0 if dtype == torch.float16 else 1.0   # fp16 gradients of a 1e-3-size loss would underflow; the chain is linear in it
    with monkeypatch.context() as mp:
        E.install(mp)
        emul = arch.Discriminator_VGG_128(in_nc=3, base_nf=int(g['cfg'][0]), input_patch_size=128)
        emul.load_state_dict(_sd(g), strict=True)
        emul.compute_dtype = dtype
        emul.train()
        e_out, e_gx, e_grads, _ = _run(emul, g, gscale)
    net = _mirror(g, dtype)
    net.train()
    out, gx, grads, running = _run(net, g, gscale)
    cos = lambda a, b: F.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()
    ref_out, ref_gx = torch.from_numpy(g['out']), torch.from_numpy(g['gx'])
    assert out.shape == (4, 1)
    assert rel_err(out, ref_out)[0] < tol_out, (out.flatten(), ref_out.flatten())
    assert rel_err(out, e_out)[0] < tol_out, (out.flatten(), e_out.flatten())
    for k in g.files:
        if k.startswith('r:'):
            assert rel_err(running[k[2:]], torch.from_numpy(g[k]))[0] < tol_out, k
    assert cos(gx, ref_gx) > min_cos and rel_err(gx, ref_gx)[1] < tol_l2, (cos(gx, ref_gx), rel_err(gx, ref_gx))
    assert rel_err(gx, e_gx)[1] < tol_same, rel_err(gx, e_gx)
    bad, worst, worst_same = [], (0.0, ''), (0.0, '')
    for k in g.files:
        if not k.startswith('g:'):
            continue
        name = k[2:]
        # a conv bias in front of a batch norm has an analytically zero gradient (round-off noise on every side)
        if name.endswith('.bias') and grads[name[:-5] + '.weight'].dim() == 4 and name != 'features.0.bias':
            continue
        a, b = grads[name], torch.from_numpy(g[k])
        el2, c, same = rel_err(a, b)[1], cos(a, b), rel_err(a, e_grads[name])[1]
        worst, worst_same = max(worst, (el2, name)), max(worst_same, (same, name))
        # BatchNorm scale / shift gradients are 8..64-element sums over all positions with heavy cancellation: twice the bound
        k1 = 2.0 if a.dim() == 1 and name.startswith('features') else 1.0
        if not (el2 < k1 * tol_l2 and c > 1 - k1 * (1 - min_cos) and same < k1 * tol_same):
            bad.append((name, round(el2, 4), round(c, 5), round(same, 5)))
    assert not bad, bad
    print('discriminator gradients (%s): worst vs reference fp32 %s rel-L2 %.2e; worst vs host arithmetic with the same operand rounding %s %.2e'
          % (dtype, worst[1], worst[0], worst_same[1], worst_same[0]))
    # eval mode: running statistics
    net.eval()
    with torch.no_grad():
        net.load_state_dict({**_sd(g), **{k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('r:')}}, strict=True)
        out_eval = net(torch.from_numpy(g['x'].astype(np.float32)).to(DEV))
    assert rel_err(out_eval.cpu(), torch.from_numpy(g['out_eval']))[0] < tol_out


def test_discriminator_reference_size_against_oracle():
    """base_nf 64 (14.5 M parameters), batch 4 of 128x128: logits and image gradient against the oracle on the host"""
    from esr_b200 import ops
    from oracle import esr_oracle as O
    import models.modules.architecture as arch
    import models.networks as networks
    ops.device_check()
    torch.manual_seed(11)
    net = arch.Discriminator_VGG_128(in_nc=3, base_nf=64)
    networks.init_weights(net, 'kaiming', scale=1)
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(p.half().float())
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.rand(4, 3, 128, 128).half().float()
    xr = x.clone().requires_grad_(True)
    ref = O.discriminator_vgg128_forward(xr, sd, training=True)
    wt = torch.tensor([[1.0], [-0.5], [0.25], [2.0]])
    (ref * wt).sum().backward()
    net = net.to(DEV)
    net.compute_dtype = torch.float16
    net.train()
    for p in net.parameters():
        p.requires_grad = False
    xd = x.to(DEV).requires_grad_(True)
    out = net(xd)
    assert rel_err(out.detach().cpu(), ref.detach())[0] < 1e-2, (out.flatten(), ref.flatten())
    (out * (16.0 * wt.to(DEV))).sum().backward()
    gx = xd.grad.cpu() / 16.0
    c = F.cosine_similarity(gx.flatten().double(), xr.grad.flatten().double(), dim=0).item()
    assert c > 0.995 and rel_err(gx, xr.grad)[1] < 0.1, (c, rel_err(gx, xr.grad))


def _gan_opt(tmp_path, **train_over):
    class ND(dict):
        def __missing__(self, k):
            return None
    train = ND(pixel_weight=1e-2, pixel_criterion='l1', gan_type='vanilla', gan_weight=5e-3, lr_G=1e-4, beta1_G=0.9, weight_decay_G=0,
               lr_D=1e-4, beta1_D=0.9, weight_decay_D=0, D_update_ratio=1, D_init_iters=0, lr_scheme='MultiStepLR', lr_steps=[1000], lr_gamma=0.5,
               grad_accumulation_steps_G=1, grad_accumulation_steps_D=1)
    train.update(train_over)
    return ND(model='srragan', scale=4, gpu_ids=[0], is_train=True, range=[0, 1], train=train,
              datasets=ND(train=ND(patch_size=144, batch_size=4)),
              path=ND(models=str(tmp_path / 'models'), pretrained_model_G=None, log=str(tmp_path)),
              network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input=None, latent_input_domain=None, latent_channels=None,
                           norm_type=None, mode='CNA', nf=32, nb=1, in_nc=3, out_nc=3, gc=32, scale=4),
              network_D=ND(which_model_D='discriminator_vgg_128', norm_type='batch', act_type='leakyrelu', mode='CNA', nf=16, in_nc=3))


def test_srragan_model_gan_training_step(tmp_path):
    """create_model -> feed_data -> optimize_parameters with the relativistic GAN branch (SRRaGAN_model.py:340-414, 466-479):
    the critic learns to separate real patches from generated ones (its loss falls, `D_logits_diff` rises), the generator
    receives a finite GAN gradient and keeps training, checkpoints of both networks round-trip."""
    from esr_b200 import ops
    from models import create_model
    ops.device_check()
    torch.manual_seed(8)
    model = create_model(_gan_opt(tmp_path, lr_D=4e-4), accumulation_steps_per_batch=1)
    netD = model.netD.module
    assert netD.classifier[0].in_features == 16 * 8 * 2 * 2          # (144 - 80) / 32 = 2
    lr = torch.rand(4, 3, 36, 36)
    hr = torch.rand(4, 3, 144, 144)                                  # independent of LR, as in the bench's synthetic pairs
    wD0 = [p.detach().clone() for p in netD.parameters()]
    wG0 = [p.detach().clone() for p in model.netG.parameters() if p.requires_grad]
    for it in range(12):
        model.feed_data({'LR': lr, 'HR': hr})
        model.optimize_parameters()
        if it == 0:   # gradient step 0: D steps, G idles (generator_step = gradient_step_num > D_init_iters)
            assert any(not torch.equal(a, p.detach()) for a, p in zip(wD0, netD.parameters()))
            assert all(torch.equal(a, p.detach()) for a, p in zip(wG0, [p for p in model.netG.parameters() if p.requires_grad]))
    d_loss = [v for _, v in model.log_dict['l_d_real_fake']]
    assert len(d_loss) == 12 and all(np.isfinite(d_loss)), d_loss
    assert np.mean(d_loss[-3:]) < d_loss[0], d_loss
    assert model.log_dict['D_logits_diff'][-1][1] > model.log_dict['D_logits_diff'][0][1]
    g_gan = [v for _, v in model.log_dict['l_g_gan']]
    assert len(g_gan) == 11 and all(np.isfinite(g_gan)), g_gan
    assert any(not torch.equal(a, p.detach()) for a, p in zip(wG0, [p for p in model.netG.parameters() if p.requires_grad]))
    assert all(torch.isfinite(p).all() for p in netD.parameters())
    assert int(netD.features[3].num_batches_tracked) > 12            # every critic call in train mode updates the running statistics
    import os
    os.makedirs(str(tmp_path / 'models'), exist_ok=True)
    model.save(12)
    assert os.path.exists(os.path.join(str(tmp_path / 'models'), '12_D.pth'))
    sd = torch.load(os.path.join(str(tmp_path / 'models'), '12_D.pth'), map_location='cpu')['model_state_dict']
    assert list(sd.keys()) == list(netD.state_dict().keys())
