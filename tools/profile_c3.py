"""One full SRRaGAN training step at the config-3 per-GPU shape inside a cudaProfilerStart/Stop range, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file <csv> python tools/profile_c3.py"""
import contextlib, io, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, 'explorable-super-resolution_b200'), REPO):
    sys.path.insert(0, p)
import torch
from models import create_model


class ND(dict):
    def __missing__(self, k):
        return None


tr = ND(pixel_weight=1e-2, pixel_criterion='l1', feature_weight=1.0, feature_criterion='l1', gan_type='vanilla', gan_weight=5e-3,
        lr_G=1e-4, beta1_G=0.9, weight_decay_G=0, lr_D=1e-4, beta1_D=0.9, weight_decay_D=0, D_update_ratio=1, D_init_iters=0,
        lr_scheme='MultiStepLR', lr_steps=[100000], lr_gamma=0.5, grad_accumulation_steps_G=1, grad_accumulation_steps_D=1)
opt = ND(model='srragan', scale=4, gpu_ids=[0], is_train=True, range=[0, 1], train=tr, datasets=ND(train=ND(patch_size=208, batch_size=4)),
         path=ND(models='/tmp/esr_prof_c3/models', pretrained_model_G=None, log='/tmp/esr_prof_c3'),
         network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input=None, latent_input_domain=None, latent_channels=None,
                      norm_type=None, mode='CNA', nf=64, nb=23, in_nc=3, out_nc=3, gc=32, scale=4),
         network_D=ND(which_model_D='discriminator_vgg_128', norm_type='batch', act_type='leakyrelu', mode='CNA', nf=64, in_nc=3))
with contextlib.redirect_stdout(io.StringIO()):
    model = create_model(opt)
lr_img, hr_img = torch.rand(4, 3, 52, 52), torch.rand(4, 3, 208, 208)
for _ in range(3):
    model.feed_data({'LR': lr_img, 'HR': hr_img})
    model.optimize_parameters()
torch.cuda.synchronize()
torch.cuda.profiler.start()
model.feed_data({'LR': lr_img, 'HR': hr_img})
model.optimize_parameters()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled one step')
