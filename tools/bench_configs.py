"""Timings of the other BASELINE.json configurations on one B200 (bench.py itself measures configs[1]):
  C1  RRDBNet nf32 nb4 x4, 1x3x128x128 forward (+ parity against the reference's golden crop)
  C3g generator side of the SRRaGAN train step: CEM(RRDB nf64 nb23), per-GPU batch 4 of 52x52 LR, pixel + VGG-feature loss, Adam
  C3  the full SRRaGAN train step of that configuration: D step (Discriminator_VGG_128 on 128x128 crops, relativistic loss, Adam) +
      G step (pixel + VGG-feature + relativistic GAN loss through the frozen critic, Adam)
  D   Discriminator_VGG_128 alone, forward + backward (parameter gradients), batches of 4 and 32 crops of 128x128
  C4  Z_optimizer, objective l1, 100 iterations over 8 regions of 64x64 LR (padded to 84x84), latent model
  C5  RRDBNet nf128 nb23 x8, 1x3x128x128 -> 1024x1024 forward and forward+backward (L1)
Prints one JSON line per configuration."""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, 'explorable-super-resolution_b200'), REPO, os.path.join(REPO, 'tests')):
    sys.path.insert(0, p)
import contextlib, io
import numpy as np
import torch
from esr_b200 import lib, ops
import models.modules.architecture as arch
from CEM.CEMnet import CEMnet, Get_CEM_Conf

dev = torch.device('cuda')
ops.device_check()


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.launch_count()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (lib.launch_count() - n0) // reps


class ND(dict):
    def __missing__(self, k):
        return None


which = sys.argv[1:] or ['C1', 'C3g', 'C4', 'C5']
if 'C1' in which:
    from util import golden
    import models.networks as networks
    g = golden('c1_seeded')
    torch.manual_seed(0)
    net = arch.RRDBNet(3, 3, 32, 4, upscale=4, num_latent_channels=0)
    with contextlib.redirect_stdout(io.StringIO()):
        networks.init_weights(net, 'kaiming', scale=0.1)
    net = net.to(dev)
    x = torch.from_numpy(g['x']).to(dev)
    with torch.no_grad():
        y = net(x)
        ms, nl = timed(lambda: net(x), reps=20)
    err = float((y[:, :, 200:264, 200:264].cpu() - torch.from_numpy(g['y_crop'])).abs().max()) / float(g['y_absmax'])
    print(json.dumps({'config': 'C1', 'ms': ms, 'HR_MP_per_s': 512 * 512 / 1e6 / (ms * 1e-3), 'launches': nl,
                      'max_err_vs_reference_rel': err}), flush=True)

if 'C3g' in which:
    from models import create_model
    train = ND(pixel_weight=1e-2, pixel_criterion='l1', feature_weight=1.0, feature_criterion='l1', lr_G=1e-4, beta1_G=0.9, weight_decay_G=0,
               lr_scheme='MultiStepLR', lr_steps=[100000], lr_gamma=0.5, grad_accumulation_steps_G=1, grad_accumulation_steps_D=1)
    opt = ND(model='srragan', scale=4, gpu_ids=[0], is_train=True, range=[0, 1], train=train, datasets=ND(train=ND(patch_size=208, batch_size=4)),
             path=ND(models='/tmp/esr_c3/models', pretrained_model_G=None, log='/tmp/esr_c3'),
             network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input=None, latent_input_domain=None, latent_channels=None,
                          norm_type=None, mode='CNA', nf=64, nb=23, in_nc=3, out_nc=3, gc=32, scale=4))
    with contextlib.redirect_stdout(io.StringIO()):
        model = create_model(opt)
    lr_img, hr_img = torch.rand(4, 3, 52, 52), torch.rand(4, 3, 208, 208)

    def step():
        model.feed_data({'LR': lr_img, 'HR': hr_img})
        model.optimize_parameters()
    ms, nl = timed(step, warm=3, reps=5)
    print(json.dumps({'config': 'C3 generator side (pixel + VGG feature loss, no discriminator), batch 4 of 52x52 LR', 'ms_per_step': ms,
                      'HR_MP_per_s': 4 * 208 * 208 / 1e6 / (ms * 1e-3), 'launches': nl, 'l_g_pix': model.log_dict['l_g_pix'][-1][1]}), flush=True)
    del model

if 'C3' in which:
    from models import create_model
    train = ND(pixel_weight=1e-2, pixel_criterion='l1', feature_weight=1.0, feature_criterion='l1', gan_type='vanilla', gan_weight=5e-3,
               lr_G=1e-4, beta1_G=0.9, weight_decay_G=0, lr_D=1e-4, beta1_D=0.9, weight_decay_D=0, D_update_ratio=1, D_init_iters=0,
               lr_scheme='MultiStepLR', lr_steps=[100000], lr_gamma=0.5, grad_accumulation_steps_G=1, grad_accumulation_steps_D=1)
    opt = ND(model='srragan', scale=4, gpu_ids=[0], is_train=True, range=[0, 1], train=train, datasets=ND(train=ND(patch_size=208, batch_size=4)),
             path=ND(models='/tmp/esr_c3/models', pretrained_model_G=None, log='/tmp/esr_c3'),
             network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input=None, latent_input_domain=None, latent_channels=None,
                          norm_type=None, mode='CNA', nf=64, nb=23, in_nc=3, out_nc=3, gc=32, scale=4),
             network_D=ND(which_model_D='discriminator_vgg_128', norm_type='batch', act_type='leakyrelu', mode='CNA', nf=64, in_nc=3))
    with contextlib.redirect_stdout(io.StringIO()):
        model = create_model(opt)
    lr_img, hr_img = torch.rand(4, 3, 52, 52), torch.rand(4, 3, 208, 208)

    def step():
        model.feed_data({'LR': lr_img, 'HR': hr_img})
        model.optimize_parameters()
    ms, nl = timed(step, warm=3, reps=5)
    print(json.dumps({'config': 'C3 full SRRaGAN step (D step + G step with pixel + VGG feature + relativistic GAN loss), batch 4 of 52x52 LR',
                      'ms_per_step': ms, 'HR_MP_per_s': 4 * 208 * 208 / 1e6 / (ms * 1e-3), 'launches': nl,
                      'l_d_real_fake': float(model.log_dict['l_d_real_fake'][-1][1]), 'l_g_gan': float(model.log_dict['l_g_gan'][-1][1]),
                      'D_logits_diff': float(model.log_dict['D_logits_diff'][-1][1])}), flush=True)
    del model

if 'D' in which:
    import models.networks as networks
    torch.manual_seed(0)
    netD = arch.Discriminator_VGG_128(3, 64)
    with contextlib.redirect_stdout(io.StringIO()):
        networks.init_weights(netD, 'kaiming', scale=1)
    netD = netD.to(dev).train()
    optD = torch.optim.Adam(netD.parameters(), lr=1e-4)
    for bsz in (4, 32):
        img = torch.rand(bsz, 3, 128, 128, device=dev)
        with torch.no_grad():
            ms_f, nl_f = timed(lambda: netD(img), warm=3, reps=10)

        def dstep():
            optD.zero_grad(set_to_none=True)
            netD(img).mean().backward()
            optD.step()
        ms_t, nl_t = timed(dstep, warm=3, reps=10)
        gf = 4.454 * bsz   # GFLOP forward, algorithmic (the space-to-depth form of the stride-2 convs executes more)
        print(json.dumps({'config': 'Discriminator_VGG_128 nf64, %d crops of 128x128' % bsz, 'fwd_ms': ms_f, 'fwd_TFLOPs_algorithmic': gf / ms_f,
                          'train_ms': ms_t, 'train_TFLOPs_algorithmic': 3 * gf / ms_t, 'launches_fwd': nl_f, 'launches_train': nl_t}), flush=True)
    del netD

if 'C4' in which:
    from models import create_model
    from Z_optimization import Z_optimizer
    opt = ND(model='srragan', scale=4, gpu_ids=[0], is_train=False, range=[0, 1], path=ND(models='/tmp/esr_c4/models', pretrained_model_G=None, log='/tmp/esr_c4'),
             network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input='all_layers', latent_input_domain='HR_downscaled',
                          latent_channels=3, norm_type=None, mode='CNA', nf=64, nb=23, in_nc=3, out_nc=3, gc=32, scale=4))
    with contextlib.redirect_stdout(io.StringIO()):
        model = create_model(opt)
    torch.manual_seed(0)
    with torch.no_grad():
        for p in model.netG.parameters():
            if p.dim() == 4 and p.shape[-1] == 3 and p.requires_grad is not None and 'Filter' not in str(p.shape):
                pass
        for n_, p in model.netG.named_parameters():
            if 'Filter_OP' not in n_:
                p.normal_(0, 0.02) if p.dim() > 1 else p.zero_()
    data = {'LR': torch.rand(8, 3, 64, 64, device=dev), 'desired': torch.rand(8, 3, 256, 256, device=dev)}
    model.feed_data({'LR': data['LR'], 'Z': 0}, need_GT=False)
    model.test()
    for iters in (5, 100):
        zo = Z_optimizer(objective='l1', Z_size=[256, 256], model=model, Z_range=1.0, max_iters=iters, data=data, initial_LR=0.1, batch_size=8)
        torch.cuda.synchronize(); n0 = lib.launch_count(); t0 = time.perf_counter()
        zo.optimize()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({'config': 'C4 Z_optimizer l1, 100 iterations, 8 regions of 64x64 LR', 's_total': dt, 'ms_per_iter': dt * 10, 'launches': lib.launch_count() - n0,
                      'loss_first': zo.loss_values[0], 'loss_last': zo.loss_values[-1]}), flush=True)
    del model

if 'C5' in which:
    torch.manual_seed(0)
    net = arch.RRDBNet(3, 3, 128, 23, upscale=8, num_latent_channels=0).to(dev)
    with torch.no_grad():
        for p in net.parameters():
            p.normal_(0, 0.01) if p.dim() > 1 else p.zero_()
    x = torch.rand(1, 3, 128, 128, device=dev)
    hr = torch.rand(1, 3, 1024, 1024, device=dev)
    for p in net.parameters():
        p.requires_grad_(False)
    with torch.no_grad():
        ms_f, nl = timed(lambda: net(x))
    for p in net.parameters():
        p.requires_grad_(True)
    optim = torch.optim.Adam(net.parameters(), lr=1e-5)

    def step():
        optim.zero_grad(set_to_none=True)
        (net(x) - hr).abs().mean().backward()
        optim.step()
    ms_t, nl_t = timed(step, warm=2, reps=3)
    fl = 1.7667e6 * 1024 * 1024
    print(json.dumps({'config': 'C5 RRDBNet nf128 nb23 x8, 1x128x128 LR per GPU', 'fwd_ms': ms_f, 'fwd_TFLOPs': fl / ms_f / 1e9, 'fwd_HR_MP_per_s': 1.048576 / (ms_f * 1e-3),
                      'train_ms': ms_t, 'train_HR_MP_per_s': 1.048576 / (ms_t * 1e-3), 'launches_fwd': nl, 'launches_train': nl_t}), flush=True)
assert lib.watchdog()[0] == 0
