"""Shared helpers for the tests: golden fixtures and mirror-network construction."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def golden_state_dict(g, prefix=''):
    return {prefix + k[2:]: torch.from_numpy(g[k].astype(np.float32)) for k in g.files if k.startswith('w:')}


def rel_err(a, b):
    """(max-abs error / max-abs reference, relative L2 error)"""
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max()).item(), ((a - b).norm() / b.norm()).item()


def mirror_rrdb(g, **extra):
    """Build the esr_b200 RRDBNet mirror for a golden fixture and load the fixture's weights."""
    import models.modules.architecture as arch
    nf, nb, s, z = [int(v) for v in g['cfg']]
    kw = dict(in_nc=3, out_nc=3, nf=nf, nb=nb, upscale=s, num_latent_channels=z)
    if z:
        kw['latent_input'] = 'all_layers_HR_downscaled'
    kw.update(extra)
    net = arch.RRDBNet(**kw)
    missing = net.load_state_dict(golden_state_dict(g), strict=True)
    return net
