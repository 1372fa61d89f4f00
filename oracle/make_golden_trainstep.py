"""Orchestration golden fixture: the UNMODIFIED reference `SRRaGANModel` (models/SRRaGAN_model.py) driven on CPU for several
`feed_data` / `optimize_parameters` calls with small STAND-IN generator / critic modules injected through `networks.define_G` /
`define_D` (so the fixture pins the training-step logic itself - D/G scheduling, loss weights, relativistic losses, gradient
accumulation, Adam steps - independently of the network kernels).  Stores the data, the initial stand-in weights, every logged
loss series and the final weights.  Build container only (`python oracle/make_golden_trainstep.py`); the fixture is committed."""
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

PATCH, SCALE, BATCH = 32, 4, 4

# the reference's load_log (SRRaGAN_model.py:653-675) reads object arrays it saved itself: numpy >= 1.16.3 needs allow_pickle for that
_np_load = np.load
np.load = lambda *a, **k: _np_load(*a, **{'allow_pickle': True, **k})


class ND(dict):
    def __missing__(self, k):
        return None


class GStand(nn.Module):
    def __init__(self, in_nc=3):
        super().__init__()
        self.c1 = nn.Conv2d(in_nc, 8, 3, padding=1)
        self.c2 = nn.Conv2d(8, 3, 3, padding=1)

    def forward(self, x):
        return self.c2(nn.functional.interpolate(nn.functional.leaky_relu(self.c1(x), 0.2), scale_factor=SCALE, mode='nearest'))


class FStand(nn.Module):
    def __init__(self):
        super().__init__()
        self.features = nn.Sequential(nn.Conv2d(3, 6, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Conv2d(6, 6, 3, padding=1))

    def forward(self, x):
        return self.features(x)


class DiscriminatorStand(nn.Module):
    def __init__(self, PATCH=PATCH):
        super().__init__()
        self.features = nn.Sequential(nn.Conv2d(3, 8, 4, stride=4), nn.LeakyReLU(0.2), nn.Conv2d(8, 8, 4, stride=4), nn.BatchNorm2d(8), nn.LeakyReLU(0.2))
        self.classifier = nn.Linear(8 * (PATCH // 16) ** 2, 1)

    def forward(self, x):
        return self.classifier(self.features(x).flatten(1))


def make_opt(tmp, variant, latent=False):
    train = ND(pixel_weight=1e-2, pixel_criterion='l1', gan_type='vanilla', gan_weight=5e-3, range_weight=0.5, lr_G=1e-3, beta1_G=0.9, weight_decay_G=0,
               lr_D=2e-3, beta1_D=0.9, weight_decay_D=0, D_update_ratio=1, D_init_iters=0, lr_scheme='MultiStepLR', lr_steps=[1000], lr_gamma=0.5,
               grad_accumulation_steps_G=1, grad_accumulation_steps_D=1, resume=0)
    train.update(variant)
    patch = 96 if latent else PATCH
    return ND(model='srragan', scale=SCALE, gpu_ids=None, is_train=True, range=[0, 1], train=train,
              datasets=ND(train=ND(patch_size=patch, batch_size=2 if latent else BATCH)),
              path=ND(models=os.path.join(tmp, 'models'), pretrained_model_G=None, pretrained_model_D=None, log=tmp, experiments_root=tmp),
              network_G=ND(which_model_G='RRDB_net', CEM_arch=0, latent_input='all_layers' if latent else 'None', latent_input_domain='HR_downscaled',
                           latent_channels='SVDinNormedOut_structure_tensor' if latent else 0,
                           norm_type=None, mode='CNA', nf=8, nb=1, in_nc=3, out_nc=3, gc=32, scale=SCALE),
              network_D=ND(which_model_D='discriminator_vgg_128', norm_type='batch', act_type='leakyrelu', mode='CNA', nf=8, in_nc=3))


VARIANTS = {
    'relativistic': dict(),
    'accumulate': dict(grad_accumulation_steps_G=2, grad_accumulation_steps_D=2),
    'plain_gan_ratio2': dict(D_update_ratio=2, _relativistic=0),
    'lsgan': dict(gan_type='lsgan'),
    'wgan_plain': dict(gan_type='wgan', _relativistic=0),
    'init_iters': dict(D_init_iters=2),
    'acc_d2_g1': dict(grad_accumulation_steps_D=2, grad_accumulation_steps_G=1),
    'no_gan': dict(gan_weight=None),
    'latent': dict(latent_weight=1.0, _latent=1),
    'lr_drop': dict(steps_4_loss_std=2, std_4_lr_drop=1e-12, lr_gamma=0.5, _loop=1),
    'optimalZ': dict(latent_weight=1.0, optimalZ_loss_type='l1', optimalZ_loss_weight=10.0, Num_Z_iterations=[10, 3], _latent=1),
    'feature': dict(feature_weight=1.0, feature_criterion='l1'),
    'feature_l2': dict(feature_weight=0.5, feature_criterion='l2', pixel_criterion='l2', gan_weight=None),
    'hinge': dict(hinge_threshold=0.05, _relativistic=0),
    'wgan_gp': dict(gan_type='wgan-gp', gp_weight=10.0, _relativistic=0),
    'verify_past': dict(D_verification='past', D_valid_Steps_4_G_update=2, min_D_prob_ratio_4_G=1.0, min_mean_D_correct=0.4, lr_D=2e-2),
    'verify_convergence': dict(D_verification='convergence', steps_4_D_convergence=3, steps_4_loss_std=3, lr_change_ratio=0.01, lr_D=2e-2),
}
N_CALLS = 8


class ValLoader:
    """what perform_validation touches of a DataLoader: iteration over batched samples and `.dataset` (un-batched samples)"""

    def __init__(self, lr, hr):
        self.dataset = [{'LR': l, 'HR': h, 'HR_path': 'img%d.png' % i} for i, (l, h) in enumerate(zip(lr, hr))]

    def __iter__(self):
        for d in self.dataset:
            yield {'LR': d['LR'].unsqueeze(0).clone(), 'HR': d['HR'].unsqueeze(0).clone(), 'HR_path': [d['HR_path']]}

    def __len__(self):
        return len(self.dataset)


def run_validation(model, tmp, data):
    from collections import OrderedDict
    model.opt['path']['val_images'] = os.path.join(tmp, 'val_images')
    os.makedirs(model.opt['path']['val_images'], exist_ok=True)
    loader = ValLoader(data['LR'][0], data['HR'][0])
    print_rlt = OrderedDict(psnr=0.0)
    model.im_collages = []
    model.gradient_step_num = 7
    sr = model.perform_validation(data_loader=loader, cur_Z=0, print_rlt=print_rlt, first_eval=True, save_images=True)
    files = sorted(os.listdir(model.opt['path']['val_images']))
    return {'val:psnr': np.array(print_rlt['psnr']), 'val:sr_mean': np.array([float(np.mean(im)) for im in sr]), 'val:collage': model.im_collages[-1],
            'val:n_files': np.array(len(files)), 'val:generator_changed': np.array(float(model.generator_changed))}


def run(model_cls, networks, tmp, variant_name, data):
    variant = dict(VARIANTS[variant_name])
    rel = variant.pop('_relativistic', None)
    latent = bool(variant.pop('_latent', 0))
    train_loop = bool(variant.pop('_loop', 0))
    opt = make_opt(tmp, variant, latent)
    patch = opt['datasets']['train']['patch_size']
    if rel is not None:
        opt['network_D']['relativistic'] = rel

    def define_G(opt, **kw):
        torch.manual_seed(100)
        return GStand(3 + 3 * SCALE ** 2 if latent else 3)

    def define_D(opt, **kw):
        torch.manual_seed(200)
        return DiscriminatorStand(patch - 80 if latent else patch)
    def define_F(opt, **kw):
        torch.manual_seed(400)
        net = FStand()
        for p_ in net.parameters():
            p_.requires_grad = False
        return net.eval()
    old = networks.define_G, networks.define_D, networks.define_F
    networks.define_G, networks.define_D, networks.define_F = define_G, define_D, define_F
    try:
        acc = max(opt['train']['grad_accumulation_steps_G'], opt['train']['grad_accumulation_steps_D'])
        model = model_cls(opt, accumulation_steps_per_batch=acc)
    finally:
        networks.define_G, networks.define_D, networks.define_F = old
    init = {'G0:' + k: v.detach().clone().numpy() for k, v in model.netG.state_dict().items()}
    if model.D_exists:
        init.update({'D0:' + k: v.detach().clone().numpy() for k, v in model.netD.state_dict().items()})
    torch.manual_seed(5)      # feed_data draws the latent codes from the global generator
    key = 'lat' if latent else ''
    lrs = []
    for it in range(N_CALLS):
        if train_loop:      # what train.py:92-101,187-189 does around the step: checkpoint + log, then the loss-driven lr rule
            model.gradient_step_num = model.step // model.max_accumulation_steps
            model.save(model.gradient_step_num)
            model.save_log()
        model.feed_data({'LR': data[key + 'LR'][it].clone(), 'HR': data[key + 'HR'][it].clone()})
        model.optimize_parameters()
        if train_loop:
            too_low = model.update_learning_rate(model.gradient_step_num)
            lrs.append([model.step, model.optimizer_G.param_groups[0]['lr'], model.optimizer_D.param_groups[0]['lr'], float(too_low)])
    logs = {'log:' + k: np.array(v, dtype=np.float64) for k, v in model.log_dict.items()
            if len(v) > 0 and k in ('l_g_pix', 'l_g_fea', 'l_g_range', 'l_g_gan', 'l_d_real', 'l_d_fake', 'D_real', 'D_fake', 'D_logits_diff', 'Correctly_distinguished',
                                    'l_g_latent_0', 'l_g_latent_1', 'l_g_latent_2', 'l_g_optimalZ', 'l_d_gp')}
    if train_loop:
        logs['log:lrs'] = np.array(lrs, dtype=np.float64)
        logs['log:D_loss_STD'] = np.array(model.log_dict['D_loss_STD'], dtype=np.float64)
        logs['log:LR_decrease_steps'] = np.array([d[0] for d in model.log_dict['LR_decrease']], dtype=np.float64)
    if variant_name == 'no_gan':      # the validation pass train.py:150-175 runs on the trained generator
        logs.update(run_validation(model, tmp, data))
    final = {'G1:' + k: v.detach().numpy() for k, v in model.netG.state_dict().items()}
    if model.D_exists:
        final.update({'D1:' + k: v.detach().numpy() for k, v in model.netD.state_dict().items()})
    return init, logs, final


def main():
    import tempfile
    import models.networks as networks
    from models.SRRaGAN_model import SRRaGANModel
    import Z_optimization as Zmod

    class TorchProxy:       # Z_optimization.py hard-codes torch.device('cuda'): redirected in memory for this CPU run
        def __getattr__(self, k):
            return getattr(torch, k)

        def device(self, *a, **k):
            return torch.device('cpu')
    Zmod.torch = TorchProxy()
    g = torch.Generator().manual_seed(31)
    q = lambda t: t.half().float()      # fp16-exact values: the fixture stores them in 16 bits
    data = {'LR': q(torch.rand(N_CALLS, BATCH, 3, PATCH // SCALE, PATCH // SCALE, generator=g)), 'HR': q(torch.rand(N_CALLS, BATCH, 3, PATCH, PATCH, generator=g))}
    data['latLR'], data['latHR'] = q(torch.rand(N_CALLS, 2, 3, 24, 24, generator=g)), q(torch.rand(N_CALLS, 2, 3, 96, 96, generator=g))
    arrays = {k: v.numpy().astype(np.float16) for k, v in data.items()}
    for name in VARIANTS:
        with tempfile.TemporaryDirectory() as tmp:
            os.makedirs(os.path.join(tmp, 'models'))
            init, logs, final = run(SRRaGANModel, networks, tmp, name, data)
        for d in (init, logs, final):
            arrays.update({name + '/' + k: v for k, v in d.items()})
        print(name, {k: v.shape for k, v in logs.items()})
    save('trainstep_orchestration', **arrays)


if __name__ == '__main__':
    main()
