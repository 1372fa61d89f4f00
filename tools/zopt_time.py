"""C4 leg alone (Z_optimizer l1, 100 iterations x 8 regions), messages visible:  python tools/zopt_time.py"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, 'explorable-super-resolution_b200'), REPO):
    sys.path.insert(0, p)
import torch
from models import create_model
from Z_optimization import Z_optimizer


class ND(dict):
    def __missing__(self, k):
        return None


dev = 'cuda'
o4 = ND(model='srragan', scale=4, gpu_ids=[0], is_train=False, range=[0, 1], path=ND(pretrained_model_G=None),
        network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input='all_layers', latent_input_domain='HR_downscaled',
                     latent_channels='SVDinNormedOut_structure_tensor', norm_type=None, mode='CNA', nf=64, nb=23, in_nc=3, out_nc=3, gc=32, scale=4))
m4 = create_model(o4)
gen = torch.Generator().manual_seed(0)
regions, iters = 8, 100
data = {'LR': torch.rand(regions, 3, 64, 64, generator=gen).to(dev), 'desired': torch.rand(regions, 3, 256, 256, generator=gen).to(dev)}
m4.feed_data({'LR': data['LR'], 'Z': 0}, need_GT=False)
m4.test()
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.time()
    zo = Z_optimizer(objective='l1', Z_size=[256, 256], model=m4, Z_range=1, max_iters=iters, data=data, initial_LR=0.1, batch_size=regions, loggers=None)
    zo.optimize()
    torch.cuda.synchronize()
    print('rep %d: %.3f s  final loss %.6f graph_ok %s' % (rep, time.time() - t0, zo.loss_values[-1], getattr(zo, '_graph_ok', None)), flush=True)
from esr_b200 import lib
print('launches', lib.load().esr_launch_count())
