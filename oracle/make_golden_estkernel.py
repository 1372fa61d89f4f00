"""Estimated-kernel golden fixture: the UNMODIFIED reference CEMnet designed around an externally supplied (non-separable)
down-scaling kernel (CEMnet.py:22-49 with upscale_kernel=ndarray, imresize_CEM.py:23-33,135-175), lower_magnitude_bound 0.1 as
SRRaGAN_model.py:54-56 sets for estimated kernels.  Stores the kernel, ds_kernel, inv_hTh and the margins."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from CEM.CEMnet import CEMnet, Get_CEM_Conf  # noqa: E402
from CEM.imresize_CEM import imresize  # noqa: E402
from make_golden import save  # noqa: E402


def aniso_kernel(n, sx, sy, theta, shift):
    y, x = np.mgrid[:n, :n].astype(np.float64)
    x, y = x - (n - 1) / 2 - shift[0], y - (n - 1) / 2 - shift[1]
    xr, yr = x * np.cos(theta) + y * np.sin(theta), -x * np.sin(theta) + y * np.cos(theta)
    k = np.exp(-0.5 * ((xr / sx) ** 2 + (yr / sy) ** 2))
    return k / k.sum()


def main():
    arrays = {}
    for tag, sf, k in (('x4', 4, aniso_kernel(21, 3.2, 1.6, 0.6, (0.7, -0.4))), ('x2', 2, aniso_kernel(13, 1.4, 0.9, -0.3, (0.0, 0.3)))):
        conf = Get_CEM_Conf(sf)
        conf.lower_magnitude_bound = 0.1
        cem = CEMnet(conf, upscale_kernel=k)
        arrays.update({tag + ':kernel': k, tag + ':ds_kernel': cem.ds_kernel, tag + ':inv_hTh': cem.inv_hTh,
                       tag + ':margins': np.array([cem.invalidity_margins_LR, cem.invalidity_margins_HR])})
        # the filters applied (reference's dense depth-wise convs, CPU): what the rank > 1 separable CUDA path must reproduce
        import torch
        gen = torch.Generator().manual_seed(sf)
        mod = cem.WrapArchitecture_PyTorch(None, None)
        xl = torch.rand(1, 3, 14, 18, generator=gen)
        gi = torch.rand(1, 3, 14 * sf, 18 * sf, generator=gen)
        with torch.no_grad():
            mod.train()
            out_train = mod([xl, gi])
            arrays.update({tag + ':x_lr': xl.numpy(), tag + ':g': gi.numpy(), tag + ':out_train': out_train.numpy(),
                           tag + ':down': mod.DownscaleOP(gi).numpy(), tag + ':inv': mod.Conv_LR_with_Inv_hTh_OP(xl).numpy(),
                           tag + ':up': mod.Upscale_OP(xl).numpy()})
        print(tag, cem.ds_kernel.shape, cem.inv_hTh.shape, cem.invalidity_margins_LR, cem.invalidity_margins_HR)
        imresize(None, [sf, sf], return_upscale_kernel=True, kernel='reset_2_default')
    save('cem_estimated_kernel', **arrays)


if __name__ == '__main__':
    main()
