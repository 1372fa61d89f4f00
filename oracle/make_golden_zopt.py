"""Z-optimisation golden fixture: the UNMODIFIED reference `Z_optimizer` (Z_optimization.py:328-797) run on CPU - its hard-coded
torch.device('cuda') is redirected in memory through a proxy of the `torch` name inside that module, nothing in the reference is
edited - around the reference's own SRRaGANModel with a small STAND-IN generator injected through networks.define_G.  Pins the
latent-exploration loop itself (objectives, Z = Z_range*tanh(.), Adam, best-iterate bookkeeping) independently of the network
kernels.  Stores inputs, stand-in weights, the loss curve and the optimised Z per objective.  Build container only."""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from make_golden import save  # noqa: E402

SCALE, H, W = 4, 12, 10


class ND(dict):
    def __missing__(self, k):
        return None


class GStand(nn.Module):
    """latent-aware stand-in: input [B, 3*16+3, h, w] (Z re-viewed + LR), output [B, 3, 4h, 4w]"""

    def __init__(self):
        super().__init__()
        self.c1 = nn.Conv2d(3 * SCALE ** 2 + 3, 12, 3, padding=1)
        self.c2 = nn.Conv2d(12, 3, 3, padding=1)

    def forward(self, x):
        return torch.sigmoid(self.c2(nn.functional.interpolate(nn.functional.leaky_relu(self.c1(x), 0.2), scale_factor=SCALE, mode='nearest')))


def make_opt(tmp):
    return ND(model='srragan', scale=SCALE, gpu_ids=None, is_train=False, range=[0, 1],
              path=ND(models=os.path.join(tmp, 'models'), pretrained_model_G=None, log=tmp),
              network_G=ND(which_model_G='RRDB_net', CEM_arch=0, latent_input='all_layers', latent_input_domain='HR_downscaled', latent_channels=3,
                           norm_type=None, mode='CNA', nf=8, nb=1, in_nc=3, out_nc=3, gc=32, scale=SCALE))


# (the reference's plain 'l1' objective with an image mask references a variable that only the 'scribble' objectives define,
#  Z_optimization.py:427 - it only runs in the training-time mode, where no mask exists: case 'l1' + training below)
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


def build_model(model_cls, networks, tmp):
    def define_G(opt, **kw):
        torch.manual_seed(300)
        return GStand()
    old = networks.define_G
    networks.define_G = define_G
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            model = model_cls(make_opt(tmp))
    finally:
        networks.define_G = old
    return model


def main():
    import tempfile
    import models.networks as networks
    from models.SRRaGAN_model import SRRaGANModel
    import Z_optimization as Zmod

    class TorchProxy:
        def __getattr__(self, k):
            return getattr(torch, k)

        def device(self, *a, **k):
            return torch.device('cpu')
    Zmod.torch = TorchProxy()
    # the scribble tool takes rgb2hsv / hsv2rgb from skimage (absent here: the standard hexcone conversion of esr_b200.colors is lent to the
    # reference) and subtracts comparison results from 1, which the torch of its day (uint8 masks) allowed
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'explorable-super-resolution_b200', 'esr_b200'))
    import colors as _colors
    Zmod.rgb2hsv, Zmod.hsv2rgb = _colors.rgb2hsv, _colors.hsv2rgb
    _rsub = torch.Tensor.__rsub__
    torch.Tensor.__rsub__ = lambda self, other: _rsub(self.to(torch.uint8) if self.dtype == torch.bool else self, other)
    g = torch.Generator().manual_seed(41)
    x_lr = torch.rand(1, 3, H, W, generator=g)
    desired = torch.rand(1, 3, SCALE * H, SCALE * W, generator=g)
    arrays = {'x_lr': x_lr.numpy(), 'desired': desired.numpy()}
    with tempfile.TemporaryDirectory() as tmp:
        model = build_model(SRRaGANModel, networks, tmp)
        arrays.update({'w:' + k: v.detach().numpy() for k, v in model.netG.state_dict().items()})
        for idx, (objective, extra, bs, iters, training) in enumerate(CASES):
            extra = dict(extra)
            mask_kw = {}
            if extra.pop('_masks', False):
                image_mask, Z_mask, labels = scribble_inputs()
                mask_kw = dict(image_mask=image_mask, Z_mask=Z_mask)
                extra['scribble_mask'] = labels
            data = {'LR': x_lr.expand(bs, -1, -1, -1).contiguous(), 'desired': desired, **extra}
            model.feed_data({'LR': data['LR'], 'Z': torch.zeros(bs, 3, SCALE * H, SCALE * W)}, need_GT=False)
            if mask_kw:               # a partial Z mask needs the current latent as the value outside the mask (GUI: initial_Z)
                mask_kw['initial_Z'] = 1 * model.GetLatent()
            if training is True:      # training-time use (SRRaGAN_model.py:108-112): no image mask, random initial Z drawn inside optimize()
                model.__dict__.pop('fake_H', None)
            else:
                model.test()
            torch.manual_seed(17 + idx)
            with contextlib.redirect_stdout(io.StringIO()):
                zo = Zmod.Z_optimizer(objective=objective, Z_size=[SCALE * H, SCALE * W], model=model, Z_range=1.0, max_iters=iters, data=data,
                                      initial_LR=0.1, batch_size=bs, HR_unpadder=(lambda t: t) if training is True else None, random_Z_inits=training == 'random',
                                      **mask_kw)
                Z = zo.optimize()
            arrays['%d:loss' % idx] = np.array([float(v) for v in zo.loss_values], dtype=np.float64)
            arrays['%d:Z' % idx] = Z.detach().cpu().numpy()
            arrays['%d:out' % idx] = model.fake_H.detach().cpu().numpy()
            print(idx, objective, bs, ['%.5f' % v for v in arrays['%d:loss' % idx]])
    save('zopt_orchestration', **arrays)


if __name__ == '__main__':
    main()
