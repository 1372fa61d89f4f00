"""CEM filters designed around an externally estimated (non-separable) kernel on the B200: the CUDA filters receive such kernels as
several separable terms (rank > 1 path of esr_cem_down / esr_cem_inv / esr_cem_up_add) and must reproduce the reference's dense
depth-wise convolutions (fixture: oracle/make_golden_estkernel.py)."""
import numpy as np
import pytest
import torch

from util import golden, rel_err

# (the host-side design is pinned on the CPU by tests/test_cem_design.py, the rank-1 kernels by tests/test_gpu_parity.py; this is the
#  rank > 1 path - confirmed on the driver's B200 at the end of round 1)
pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('tag,s', [('x4', 4), ('x2', 2)])
def test_estimated_kernel_filters_match_reference(tag, s):
    from CEM.CEMnet import CEMnet, Get_CEM_Conf
    from CEM.imresize_CEM import imresize
    g = golden('cem_estimated_kernel')
    conf = Get_CEM_Conf(s)
    conf.lower_magnitude_bound = 0.1
    try:
        mod = CEMnet(conf, upscale_kernel=g[tag + ':kernel']).WrapArchitecture_PyTorch(None, None).to(DEV)
        x, gi = torch.from_numpy(g[tag + ':x_lr']).to(DEV), torch.from_numpy(g[tag + ':g']).to(DEV)
        T = lambda k: torch.from_numpy(g[tag + ':' + k])
        with torch.no_grad():
            assert rel_err(mod.DownscaleOP(gi).cpu(), T('down'))[0] < 1e-5
            assert rel_err(mod.Conv_LR_with_Inv_hTh_OP(x).cpu(), T('inv'))[0] < 1e-5
            assert rel_err(mod.Upscale_OP(x).cpu(), T('up'))[0] < 1e-5
            mod.train()
            assert rel_err(mod([x, gi]).cpu(), T('out_train'))[0] < 1e-5
    finally:
        imresize(None, [s, s], return_upscale_kernel=True, kernel='reset_2_default')
