"""Parity arithmetic mode (esr_b200.precision 'parity' = esr_dtype ESR_BF16X3, bf16 hi/lo split operands, fp32 accumulation) against the
UNMODIFIED reference's fp32 results, held to north_star's tolerance: 1e-3 relative.

  * single convs (both tcgen05 kernels) and the wgrad kernel against fp64 torch on full-precision (not 16-bit-rounded) operands;
  * the generator's training backward (every weight gradient) on the reference's kink-free fixtures at 1e-3 - the test that runs at
    6e-2 in the bf16 throughput mode (tests/test_train_gpu.py);
  * BASELINE config 2's generator at FULL DEPTH (nf=64, nb=23, behind the CEM) and config 5's at FULL WIDTH (nf=128, x8): forward,
    input gradient and weight gradients against fixtures of oracle/make_golden_fulldepth.py, with the throughput modes measured
    beside them (printed, held to their own looser bounds);
  * Discriminator_VGG_128 (logits, parameter and image gradients) at 1e-3 on a kink-free fixture of the reference, VGG19 features at
    1e-3 (its image gradient passes 15 ReLU and 4 max-pool decisions: reported next to torch's own fp32-vs-fp64 disagreement)."""
import contextlib
import io

import numpy as np
import pytest
import torch

from util import golden, golden_state_dict, mirror_rrdb, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 1e-3


@pytest.fixture(autouse=True)
def _watchdog():
    yield
    from esr_b200 import lib
    wd = lib.watchdog()
    assert wd[0] == 0, 'pipeline watchdog fired: %r' % (wd,)


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()


# ------------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize('rows', ['force', False])
@pytest.mark.parametrize('cin,cout,h,w', [(64, 32, 24, 130), (96, 64, 9, 140), (19, 7, 17, 33)])
def test_split_conv_matches_fp64(rows, cin, cout, h, w):
    """x*w = x_hi*w_hi + x_lo*w_hi + x_hi*w_lo: 2^-17 per operand.  Operands are full fp32 values (NOT rounded to 16 bits)."""
    from esr_b200 import ops
    ops.device_check()
    g = torch.Generator().manual_seed(cin * 7 + cout)
    n = 2
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    b = torch.randn(cout, generator=g)
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x.double(), wt.double(), b.double(), padding=1), 0.2)
    pc = ops.PackedConv(wt.to(DEV), b.to(DEV), dtype=ops.SPLIT)
    if rows == 'force' and pc.wrows is None:
        pytest.skip('weights do not fit the row kernel')
    x16, _ = ops.pack_nchw(x.to(DEV), dtype=ops.SPLIT)
    assert x16.shape[1] == 2 * ops.planes_for(cin)
    # (a) fp32 NCHW output; (b) split 16-bit output read back as hi + lo
    out = torch.empty(n, cout, h, w, device=DEV)
    ops.conv3x3(x16, pc, lrelu=True, slope=0.2, out_nchw=out, rows=rows)
    emax, el2 = rel_err(out.cpu(), ref)
    assert emax < 3e-5 and el2 < 2e-5, (emax, el2)
    o16 = ops.alloc16(ops.SPLIT, n, ops.planes_for(cout), h, w, DEV, zero=True)
    ops.conv3x3(x16, pc, lrelu=True, slope=0.2, out16=o16, rows=rows)
    back = ops.unpack_planes(o16, cout, split=True)
    emax, el2 = rel_err(back.cpu(), ref)
    assert emax < 5e-5 and el2 < 3e-5, (emax, el2)
    # the same launch in bf16 is two decades worse: the split is doing the work
    pcb = ops.PackedConv(wt.to(DEV), b.to(DEV), dtype=torch.bfloat16)
    xb, _ = ops.pack_nchw(x.to(DEV), dtype=torch.bfloat16)
    outb = torch.empty(n, cout, h, w, device=DEV)
    ops.conv3x3(xb, pcb, lrelu=True, slope=0.2, out_nchw=outb, rows=rows)
    assert rel_err(outb.cpu(), ref)[1] > 30 * el2


def test_split_wgrad_matches_fp64():
    from esr_b200 import ops
    ops.device_check()
    g = torch.Generator().manual_seed(5)
    n, cin, cout, h, w = 2, 40, 32, 21, 70
    x = torch.randn(n, cin, h, w, generator=g)
    gy = torch.randn(n, cout, h, w, generator=g) * 1e-4            # gradient-sized values: no underflow in split bf16
    xr = x.double().requires_grad_(False)
    wd = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.conv2d(xr, wd, padding=1).backward(gy.double())
    x16, _ = ops.pack_nchw(x.to(DEV), dtype=ops.SPLIT)
    g16, _ = ops.pack_nchw(gy.to(DEV), dtype=ops.SPLIT)
    dw, db = ops.conv3x3_wgrad(x16, g16, cout, cin, split=True)
    emax, el2 = rel_err(dw.cpu(), wd.grad)
    assert emax < 5e-5 and el2 < 3e-5, (emax, el2)
    assert rel_err(db.cpu(), gy.double().sum(dim=(0, 2, 3)))[0] < 5e-5


# ------------------------------------------------------------------------------------------------ training backward, 1e-3
@pytest.mark.parametrize('tag', ['plain', 'latent'])
def test_weight_gradients_parity_mode_1e3(tag):
    from esr_b200 import ops, precision
    from CEM.CEMnet import CEMnet, Get_CEM_Conf
    ops.device_check()
    g = golden('wgrad_kinkfree_%s_train' % tag)
    net = mirror_rrdb(g)
    wrapped = CEMnet(Get_CEM_Conf(4)).WrapArchitecture_PyTorch(net, None).to(DEV)
    wrapped.train()
    x = torch.from_numpy(g['x']).to(DEV)
    with precision.use('parity'):
        out = wrapped(x)
        ref_out = torch.from_numpy(g['out'])
        assert (out.detach().cpu() - ref_out).abs().max().item() < 1e-4 * max(1.0, ref_out.abs().max().item())
        (out * torch.from_numpy(g['wt']).to(DEV)).sum().backward()
    worst, bad = (0.0, 0.0, ''), []
    for name, p in net.named_parameters():
        ref = torch.from_numpy(g['g:' + name])
        emax, el2 = rel_err(p.grad.cpu(), ref)
        worst = max(worst, (el2, emax, name))
        if not (el2 < TOL and emax < TOL):
            bad.append((name, round(emax, 6), round(el2, 6)))
    print('parity mode, worst weight gradient: %s rel-L2 %.2e max %.2e' % (worst[2], worst[0], worst[1]))
    assert not bad, bad


def test_input_gradient_parity_mode_1e3():
    """Z-optimisation's backward (frozen generator, eval-mode CEM with padding) on the kink-free latent fixture"""
    from esr_b200 import ops, precision
    from CEM.CEMnet import CEMnet, Get_CEM_Conf
    ops.device_check()
    g = golden('grad_kinkfree_latent_eval')
    net = mirror_rrdb(g)
    wrapped = CEMnet(Get_CEM_Conf(4)).WrapArchitecture_PyTorch(net, None).to(DEV)
    for p in wrapped.parameters():
        p.requires_grad_(False)
    wrapped.eval()
    x = torch.from_numpy(g['x']).to(DEV).requires_grad_(True)
    with precision.use('parity'):
        out = wrapped(x)
        (out * torch.from_numpy(g['wt']).to(DEV)).sum().backward()
    zc = x.shape[1] - 3
    emax, el2 = rel_err(x.grad[:, :zc].cpu(), torch.from_numpy(g['gx'])[:, :zc])
    print('parity mode, latent gradient: max %.2e rel-L2 %.2e' % (emax, el2))
    assert emax < TOL and el2 < TOL
    assert (out.detach().cpu() - torch.from_numpy(g['out'])).abs().max().item() < 1e-4


# ------------------------------------------------------------------------------------------------ full depth / full width
def _seeded(name):
    import models.modules.architecture as arch
    import models.networks as networks
    g = golden(name)
    nf, nb, scale, lr_hw, seed, with_cem, init100, kink_free = [int(v) for v in g['cfg']]
    torch.manual_seed(seed)
    net = arch.RRDBNet(3, 3, nf, nb, upscale=scale, num_latent_channels=0)
    with contextlib.redirect_stdout(io.StringIO()):
        networks.init_weights(net, 'kaiming', scale=init100 / 100.0)
    if kink_free:
        from oracle.fixture_util import kink_free_biases
        kink_free_biases(net, seed)
    wsum = np.array([float(v.double().sum()) for v in net.state_dict().values()])
    assert np.allclose(wsum, g['wsum'], rtol=0, atol=1e-8), 'seeded init differs from the reference init'
    x = torch.rand(1, 3, lr_hw, lr_hw, generator=torch.Generator().manual_seed(seed + 100))
    assert torch.equal(x, torch.from_numpy(g['x']))
    model = net
    if with_cem:
        from CEM.CEMnet import CEMnet, Get_CEM_Conf
        model = CEMnet(Get_CEM_Conf(scale)).WrapArchitecture_PyTorch(net, None)
        model.train()
    H = lr_hw * scale
    wt = torch.randn((1, 3, H, H), generator=torch.Generator().manual_seed(seed + 200))
    return g, net, model.to(DEV), x, wt


def _check_fulldepth(name, mode, fwd_tol, grad_tol):
    from esr_b200 import ops, precision
    ops.device_check()
    g, net, model, x, wt = _seeded(name)
    scale = float(g['y_absmax'])
    xd = x.to(DEV).requires_grad_(True)
    with precision.use(mode):
        out = model(xd)
        (out * wt.to(DEV)).sum().backward()
    y = out.detach().cpu()
    H = y.shape[-1]
    c0 = H // 2 - 32
    e_crop = (y[:, :, c0:c0 + 64, c0:c0 + 64] - torch.from_numpy(g['y_crop'])).abs().max().item() / scale
    e_rows = (y[0, :, ::32, :] - torch.from_numpy(g['y_rows'])).abs().max().item() / scale
    e_mean = abs(float(y.double().mean()) - float(g['y_mean'])) / scale
    gx_max, gx_l2 = rel_err(xd.grad.cpu(), torch.from_numpy(g['gx']))
    sd = dict(net.named_parameters())
    worst = (0.0, 0.0, '')
    for key in [k for k in g.files if k.startswith('dw:') or k.startswith('db:')]:
        pname = key[3:] + ('.weight' if key.startswith('dw:') else '.bias')
        ref = torch.from_numpy(g[key])
        if float(ref.abs().max()) == 0:
            continue
        emax, el2 = rel_err(sd[pname].grad.cpu(), ref)
        worst = max(worst, (el2, emax, pname))
    gnorm = np.array([float(p.grad.double().norm()) for p in net.parameters()])
    ref_norm = g['gnorm']
    e_norm = float(np.max(np.abs(gnorm - ref_norm) / np.maximum(ref_norm, 1e-30 + 1e-6 * ref_norm.max())))
    print('%s [%s]: forward crop %.2e rows %.2e mean %.2e of range | input grad max %.2e L2 %.2e | worst probed weight grad %s L2 %.2e max %.2e | '
          'worst |grad| norm dev %.2e' % (name, mode, e_crop, e_rows, e_mean, gx_max, gx_l2, worst[2], worst[0], worst[1], e_norm))
    assert max(e_crop, e_rows) < fwd_tol, (e_crop, e_rows)
    assert gx_l2 < grad_tol and worst[0] < grad_tol and e_norm < grad_tol, (gx_l2, worst, e_norm)
    return e_crop, gx_l2, worst[0]


def test_c2_full_depth_kinkfree_parity_mode_1e3():
    """RRDBNet nf=64 nb=23 x4 + CEM (the headline's generator) vs the unmodified reference: forward AND backward (input gradient,
    probed weight gradients, every parameter's gradient norm) within 1e-3, on the fixture whose LeakyReLU inputs stay away from 0"""
    _check_fulldepth('c2_depth_kf', 'parity', TOL, TOL)


def test_c5_full_width_kinkfree_parity_mode_1e3():
    """RRDBNet nf=128 nb=23 x8 (config 5's generator) vs the unmodified reference, forward and backward within 1e-3"""
    _check_fulldepth('c5_width_kf', 'parity', TOL, TOL)


def test_c2_full_depth_training_init_parity_mode():
    """the reference's own training init (kaiming x0.1, zero biases: pre-activations centred on the LeakyReLU kink).  The FORWARD is
    held to 1e-3; gradients are reported and held to the kink-flip bound (oracle/fixture_util.kink_free_biases explains why no
    non-bit-identical implementation can do better on such a fixture)"""
    _check_fulldepth('c2_depth', 'parity', TOL, 3e-2)


@pytest.mark.parametrize('name', ['c2_depth_kf', 'c5_width_kf', 'c2_depth'])
def test_full_depth_throughput_mode_reported(name):
    """the same fixtures through the throughput arithmetic (bf16 operands once a parameter wants a gradient): measured and printed
    beside the parity mode; held to the bound that mode is documented with"""
    _check_fulldepth(name, 'throughput', 3e-2, 2e-1 if name.endswith('_kf') else 6e-1)


def test_c2_full_depth_inference_fp16():
    """no-grad forward of the full-depth generator in the fp16 inference mode (bench `forward` leg arithmetic)"""
    from esr_b200 import ops
    ops.device_check()
    g, net, model, x, wt = _seeded('c2_depth')
    with torch.no_grad():
        y = model(x.to(DEV)).cpu()
    scale = float(g['y_absmax'])
    c0 = y.shape[-1] // 2 - 32
    e = (y[:, :, c0:c0 + 64, c0:c0 + 64] - torch.from_numpy(g['y_crop'])).abs().max().item() / scale
    print('c2_depth fp16 inference: max err %.2e of range' % e)
    assert e < 2e-3


# ------------------------------------------------------------------------------------------------ critic and VGG19
def _critic_run(fixture, mode):
    from esr_b200 import ops, precision
    import models.modules.architecture as arch
    ops.device_check()
    g = golden(fixture)
    netD = arch.Discriminator_VGG_128(3, int(g['cfg'][0]), input_patch_size=int(g['x'].shape[-1]))
    sd = {}
    for k in g.files:
        if k.startswith('w:'):
            v = torch.from_numpy(g[k])
            sd[k[2:]] = v.float() if v.dtype.is_floating_point else v
    netD.load_state_dict(sd, strict=True)
    netD = netD.to(DEV).train()
    x = torch.from_numpy(g['x'].astype(np.float32)).to(DEV).requires_grad_(True)
    with precision.use(mode):
        logits = netD(x)
        (logits * torch.from_numpy(g['wt']).to(DEV)).sum().backward()
    ref_logits = torch.from_numpy(g['out'])
    e_log = (logits.detach().cpu() - ref_logits).abs().max().item() / ref_logits.abs().max().item()
    gx_max, gx_l2 = rel_err(x.grad.cpu(), torch.from_numpy(g['gx']))
    worst = (0.0, 0.0, '')
    mods = dict(netD.named_modules())
    for name, p in netD.named_parameters():
        ref = torch.from_numpy(g['g:' + name])
        parts = name.split('.')
        if parts[0] == 'features' and parts[2] == 'bias' and isinstance(mods.get('features.%d' % (int(parts[1]) + 1)), torch.nn.BatchNorm2d):
            # a conv bias in front of a BatchNorm: the normalisation removes it, its true gradient is exactly 0 and both sides hold
            # rounding noise - compare against the scale of the same conv's weight gradient instead of relatively
            wscale = float(torch.from_numpy(g['g:features.%s.weight' % parts[1]]).abs().max())
            ztol = 1e-3 if mode == 'parity' else 5e-2
            assert float(p.grad.abs().max()) < ztol * wscale and float(ref.abs().max()) < 1e-3 * wscale, name
            continue
        emax, el2 = rel_err(p.grad.cpu(), ref)
        worst = max(worst, (el2, emax, name))
    print('critic %s [%s]: logits %.2e | image gradient max %.2e L2 %.2e | worst parameter gradient %s L2 %.2e max %.2e'
          % (fixture, mode, e_log, gx_max, gx_l2, worst[2], worst[0], worst[1]))
    return e_log, gx_l2, worst[0]


def test_discriminator_kinkfree_parity_mode_1e3():
    """Discriminator_VGG_128 vs the unmodified reference (train-mode BatchNorm): logits, image gradient and EVERY parameter gradient
    within 1e-3 on the kink-free fixture (oracle/make_golden_disc_kf.py)"""
    e_log, gx_l2, worst = _critic_run('disc_vgg128_nf8_kf', 'parity')
    assert e_log < TOL and gx_l2 < TOL and worst < TOL, (e_log, gx_l2, worst)


def test_discriminator_plain_fixture_parity_mode():
    """the round-1 fixture (pre-activations crossing the LeakyReLU kink everywhere): logits within 1e-3; gradients reported, held to
    the kink-flip bound"""
    e_log, gx_l2, worst = _critic_run('disc_vgg128_nf8', 'parity')
    assert e_log < TOL, e_log
    assert gx_l2 < 6e-2 and worst < 6e-2, (gx_l2, worst)


def test_discriminator_throughput_mode_reported():
    e_log, gx_l2, worst = _critic_run('disc_vgg128_nf8_kf', 'throughput')
    assert e_log < 5e-2 and gx_l2 < 0.3 and worst < 0.3


def test_vgg_parity_mode_1e3():
    """VGG19[:35] features and image gradient vs the same nn.Sequential in fp64"""
    from esr_b200 import ops, precision
    import models.modules.architecture as arch
    ops.device_check()
    torch.manual_seed(4)
    netF = arch.VGGFeatureExtractor(feature_layer=34, arch_config='untrained').to(DEV).eval()
    x = torch.rand(2, 3, 64, 64)
    import copy
    ref_net = torch.nn.Sequential(*[torch.nn.ReLU(inplace=False) if isinstance(m, torch.nn.ReLU) else copy.deepcopy(m)
                                    for m in netF.features.children()]).double().cpu()
    xr = x.double().requires_grad_(True)
    mean, std = netF.mean.double().cpu(), netF.std.double().cpu()
    fr = ref_net((xr - mean) / std)
    wt = torch.randn(fr.shape, generator=torch.Generator().manual_seed(9), dtype=torch.float64)
    (fr * wt).sum().backward()
    xd = x.to(DEV).requires_grad_(True)
    with precision.use('parity'):
        f = netF(xd)
        (f * wt.float().to(DEV)).sum().backward()
    f_max, f_l2 = rel_err(f.detach().cpu(), fr.detach())
    g_max, g_l2 = rel_err(xd.grad.cpu(), xr.grad)
    # context: torch's own fp32 autograd against the fp64 one on the same fixture (15 ReLU + 4 max-pool decisions flip under ANY rounding)
    x32 = x.clone().requires_grad_(True)
    f32 = copy.deepcopy(ref_net).float()((x32 - mean.float()) / std.float())
    (f32 * wt.float()).sum().backward()
    g32_l2 = rel_err(x32.grad, xr.grad)[1]
    print('VGG19, parity mode: features max %.2e L2 %.2e | image gradient max %.2e L2 %.2e cos %.6f (torch fp32 vs fp64 autograd: L2 %.2e)'
          % (f_max, f_l2, g_max, g_l2, _cos(xd.grad.cpu(), xr.grad), g32_l2))
    assert f_l2 < TOL and f_max < TOL, (f_max, f_l2)
    assert _cos(xd.grad.cpu(), xr.grad) > 0.999 and g_l2 < 5e-2, g_l2
