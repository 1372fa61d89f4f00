"""Soft histogram / dictionary objective on the B200 (esr_soft_hist_fwd / _bwd through Z_optimization.SoftHistogramLoss): the kernels
against the oracle's materialised-tensor restatement (grey levels and 3x3 / 6x6 patches), the Module against the unmodified reference's
fixture, and the `hist` / local-STD / VGG objectives of Z_optimizer stepping."""
import contextlib
import io

import numpy as np
import pytest
import torch

from util import golden, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('D,P,B', [(1, 700, 256), (9, 333, 77), (36, 150, 40)])
@pytest.mark.parametrize('dictionary', [False, True])
def test_soft_hist_kernels_match_oracle(D, P, B, dictionary):
    from oracle import esr_oracle as O
    from Z_optimization import _SoftHistFn
    g = torch.Generator().manual_seed(D + P)
    x = torch.rand(D, P, generator=g, dtype=torch.float64)
    bins = torch.rand(D, B, generator=g, dtype=torch.float64)
    T = 0.02 if D == 1 else 0.05
    xr = x.clone().requires_grad_(True)
    ref = O.soft_histogram(xr, bins, 1.0, T, dictionary)
    wt = torch.rand(ref.shape, generator=g, dtype=torch.float64)
    (ref * wt).sum().backward()
    xd = x.to(DEV).requires_grad_(True)
    out = _SoftHistFn.apply(xd, bins.to(DEV), 1.0, 1e-7, T, dictionary)
    (out * wt.to(DEV)).sum().backward()
    assert rel_err(out.detach().cpu(), ref.detach())[0] < 1e-9
    assert rel_err(xd.grad.cpu(), xr.grad)[0] < 1e-8


def test_soft_histogram_loss_matches_reference_fixture():
    from Z_optimization import SoftHistogramLoss
    g = golden('soft_hist')
    des = torch.from_numpy(g['desired']).to(DEV)
    mask = torch.from_numpy(g['image_mask']).to(DEV)
    for name, kw in (('hist', dict(temperature=5e-4, dictionary_not_histogram=False)), ('dict', dict(temperature=1e-3, dictionary_not_histogram=True)),
                     ('hist_warm', dict(temperature=2e-2, dictionary_not_histogram=False))):
        cur = torch.from_numpy(g['cur']).to(DEV).requires_grad_(True)
        mod = SoftHistogramLoss(bins=256, min=0, max=1, desired_hist_image=[des], desired_hist_image_mask=[g['desired_mask']], input_im_HR_mask=mask,
                                gray_scale=True, patch_size=1, **kw)
        loss = mod(cur)
        loss.mean().backward()
        assert np.allclose(loss.detach().cpu().numpy(), g[name + ':loss'], rtol=2e-5, atol=1e-9), (name, loss, g[name + ':loss'])
        ref = torch.from_numpy(g[name + ':grad'])
        assert float((cur.grad.cpu() - ref).abs().max()) < 2e-5 * float(ref.abs().max()), name


def _gui_model(init_Fnet=False):
    from models import create_model

    class ND(dict):
        def __missing__(self, k):
            return None
    opt = ND(model='srragan', scale=4, gpu_ids=[0], is_train=False, range=[0, 1], path=ND(pretrained_model_G=None),
             datasets=ND(train=ND(patch_size=128)),
             network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input='all_layers', latent_input_domain='HR_downscaled',
                          latent_channels='SVDinNormedOut_structure_tensor', norm_type=None, mode='CNA', nf=32, nb=1, in_nc=3, out_nc=3, gc=32, scale=4))
    torch.manual_seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        model = create_model(opt, init_Fnet=init_Fnet)
    with torch.no_grad():
        for n_, p in model.netG.named_parameters():
            if 'Filter_OP' not in n_:
                p.normal_(0, 0.05) if p.dim() > 1 else p.zero_()
    return model


@pytest.mark.parametrize('objective,extra', [('hist', {}), ('dict', {}), ('hist_patch_noDC', {}), ('local_STD_increase', {'STD_increment': 0.02}),
                                             ('local_Mag_increase', {'STD_increment': 0.02}), ('VGG', {})])
def test_z_optimizer_new_objectives_step(objective, extra):
    """the GUI's imprinting / local-contrast / perceptual tools: the loop runs on the CUDA path, losses are finite and go down"""
    from Z_optimization import Z_optimizer
    model = _gui_model(init_Fnet=objective == 'VGG')
    g = torch.Generator().manual_seed(9)
    lr = torch.rand(1, 3, 16, 16, generator=g).to(DEV)
    H = 64
    model.feed_data({'LR': lr, 'Z': 0}, need_GT=False)
    model.test()
    image_mask = np.zeros((H, H), np.float32)
    image_mask[8:56, 8:56] = 1
    z_mask = image_mask.copy()
    data = {'LR': lr, 'desired': torch.rand(1, 3, H, H, generator=g).to(DEV) * 0.5 + 0.25, 'Desired_Im_Mask': [np.ones((H, H), bool)], **extra}
    if 'hist' in objective or 'dict' in objective:
        data['desired'] = [data['desired']]
    with contextlib.redirect_stdout(io.StringIO()):
        zo = Z_optimizer(objective=objective, Z_size=[H, H], model=model, Z_range=1.0, max_iters=8, data=data, initial_LR=0.05, batch_size=1,
                         image_mask=image_mask, Z_mask=z_mask, initial_Z=model.GetLatent())
        zo.optimize()
    vals = zo.loss_values
    assert len(vals) >= 1 and all(np.isfinite(v) for v in vals), vals
    assert min(vals) <= vals[0]
