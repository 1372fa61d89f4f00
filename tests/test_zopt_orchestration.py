"""The latent-exploration loop (Z_optimization.Z_optimizer) against the UNMODIFIED reference's (oracle/make_golden_zopt.py): both
run on the CPU around their own SRRaGANModel with the same stand-in generator injected through networks.define_G, so the
comparison pins the loop itself - objectives (STD / TV / l1), Z = Z_range*tanh(.), Adam, the best-iterate bookkeeping and the
number of iterations actually run - independently of the CUDA kernels."""
import contextlib
import io

import numpy as np
import pytest
import torch
import torch.nn as nn

from util import golden

SCALE, H, W = 4, 12, 10


class ND(dict):
    def __missing__(self, k):
        return None


class GStand(nn.Module):
    def __init__(self):
        super().__init__()
        self.c1 = nn.Conv2d(3 * SCALE ** 2 + 3, 12, 3, padding=1)
        self.c2 = nn.Conv2d(12, 3, 3, padding=1)

    def forward(self, x):
        return torch.sigmoid(self.c2(nn.functional.interpolate(nn.functional.leaky_relu(self.c1(x), 0.2), scale_factor=SCALE, mode='nearest')))


CASES = [('max_STD', {}, 1, 6, False), ('min_STD', {}, 1, 6, False), ('TV', {}, 1, 6, False), ('STD_increase', {'STD_increment': 0.01}, 1, 6, False),
         ('STD_decrease', {'STD_increment': 0.02}, 1, 6, False), ('l1', {}, 2, 8, True),
         ('random_l1', {}, 3, 5, 'random'), ('max_STD', {}, 1, -4, False), ('min_STD', {}, 1, -2, False),
         ('periodicity', {'periodicity_points': [[3, 2]]}, 1, 6, False),
         ('nonInt_periodicity', {'periodicity_points': [[3.4, 1.7], [-2.2, 4.1]]}, 1, 6, False),
         ('nonInt_periodicity_Plus', {'periodicity_points': [[2.5, 3.3]], 'STD_increment': 0.01}, 1, 5, False),
         ('scribble', {'_masks': True, 'brightness_factor': 0.3}, 1, 6, False)]


def scribble_inputs():
    """region masks and scribble labels of the 'scribble' case: 1 colour, 2 brighten, 3 darken, 4 / 5 two smoothing regions"""
    image_mask = np.zeros((SCALE * H, SCALE * W), dtype=np.float32)
    image_mask[6:42, 4:36] = 1
    labels = np.zeros((SCALE * H, SCALE * W), dtype=np.int64)
    labels[8:14, 6:20] = 1
    labels[16:22, 8:18] = 2
    labels[16:22, 22:32] = 3
    labels[26:34, 6:16] = 4
    labels[28:38, 20:34] = 5
    labels[2:6, 2:10] = 1          # outside the region mask: must not count
    return image_mask, 1 * image_mask, labels


def _opt(tmp_path):
    return ND(model='srragan', scale=SCALE, gpu_ids=None, is_train=False, range=[0, 1],
              path=ND(models=str(tmp_path / 'models'), pretrained_model_G=None, log=str(tmp_path)),
              network_G=ND(which_model_G='RRDB_net', CEM_arch=0, latent_input='all_layers', latent_input_domain='HR_downscaled', latent_channels=3,
                           norm_type=None, mode='CNA', nf=8, nb=1, in_nc=3, out_nc=3, gc=32, scale=SCALE))


@pytest.mark.parametrize('idx', range(len(CASES)))
def test_z_optimizer_loop_matches_reference(monkeypatch, tmp_path, idx):
    if torch.cuda.is_available():
        pytest.skip('CPU-suite test: the stand-in generator and the fixture live on the host')
    import models.networks as networks
    import Z_optimization as Zmod
    from models.SRRaGAN_model import SRRaGANModel
    g = golden('zopt_orchestration')

    def define_G(opt, **kw):
        torch.manual_seed(300)
        return GStand()
    monkeypatch.setattr(networks, 'define_G', define_G)
    monkeypatch.setattr(Zmod, '_dev', lambda: torch.device('cpu'))
    with contextlib.redirect_stdout(io.StringIO()):
        model = SRRaGANModel(_opt(tmp_path))
    for k, v in model.netG.state_dict().items():
        assert np.array_equal(v.numpy(), g['w:' + k]), k
    objective, extra, bs, iters, training = CASES[idx]
    x_lr, desired = torch.from_numpy(g['x_lr']), torch.from_numpy(g['desired'])
    extra = dict(extra)
    mask_kw = {}
    if extra.pop('_masks', False):
        image_mask, Z_mask, labels = scribble_inputs()
        mask_kw = dict(image_mask=image_mask, Z_mask=Z_mask)
        extra['scribble_mask'] = labels
    data = {'LR': x_lr.expand(bs, -1, -1, -1).contiguous(), 'desired': desired, **extra}
    model.feed_data({'LR': data['LR'], 'Z': torch.zeros(bs, 3, SCALE * H, SCALE * W)}, need_GT=False)
    if mask_kw:
        mask_kw['initial_Z'] = 1 * model.GetLatent()
    if training is True:
        model.__dict__.pop('fake_H', None)
    else:
        model.test()
    torch.manual_seed(17 + idx)
    with contextlib.redirect_stdout(io.StringIO()):
        zo = Zmod.Z_optimizer(objective=objective, Z_size=[SCALE * H, SCALE * W], model=model, Z_range=1.0, max_iters=iters, data=data,
                              initial_LR=0.1, batch_size=bs, HR_unpadder=(lambda t: t) if training is True else None, random_Z_inits=training == 'random',
                              **mask_kw)
        Z = zo.optimize()
    ref_loss = g['%d:loss' % idx]
    own_loss = np.array([float(v) for v in zo.loss_values])
    assert own_loss.shape == ref_loss.shape, (own_loss, ref_loss)          # same number of iterations actually run
    assert np.allclose(own_loss, ref_loss, rtol=1e-3, atol=1e-6), (own_loss, ref_loss)
    assert np.allclose(Z.detach().numpy(), g['%d:Z' % idx], atol=2e-3), np.abs(Z.detach().numpy() - g['%d:Z' % idx]).max()
    assert np.allclose(model.fake_H.detach().numpy(), g['%d:out' % idx], atol=1e-4)
