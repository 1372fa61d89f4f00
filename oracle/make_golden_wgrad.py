"""Weight-gradient golden fixture: runs the UNMODIFIED reference (CPU fp32 autograd) on a CEM-wrapped RRDBNet in train
mode, loss = sum(out * Wt), and stores d(loss)/d(parameter).  Build container only (`python oracle/make_golden_wgrad.py`);
the fixture is committed.  Kink-free construction as in make_golden.py E3 (every LeakyReLU input far from 0), so the
gradient is comparable element-wise at operand precision.  With and without latent input."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import torch  # noqa: E402
from CEM.CEMnet import CEMnet, Get_CEM_Conf  # noqa: E402
from make_golden import build_rrdb, q16, save, sd_np  # noqa: E402


def main():
    g = torch.Generator().manual_seed(4321)
    for tag, z in (('plain', 0), ('latent', 3)):
        kw = dict(in_nc=3, out_nc=3, nf=32, nb=1, upscale=4, num_latent_channels=z)
        if z:
            kw['latent_input'] = 'all_layers_HR_downscaled'
        net = build_rrdb(11 + z, scale=0.1, **kw)
        with torch.no_grad():
            for name, p in net.named_parameters():
                if name.endswith('bias'):
                    sign = torch.where(torch.rand(p.shape, generator=g) < 0.5, -1.0, 1.0)
                    p.copy_(q16(sign * (2.0 + torch.rand(p.shape, generator=g))))
        margins = []
        hooks = [m.register_forward_pre_hook(lambda mod, inp: margins.append(float(inp[0].abs().min())))
                 for m in net.modules() if isinstance(m, torch.nn.LeakyReLU)]
        cem = CEMnet(Get_CEM_Conf(4))
        wrapped = cem.WrapArchitecture_PyTorch(net, None)
        wrapped.train()
        h, w = 20, 16
        lr = q16(torch.rand(2, 3, h, w, generator=g))
        if z:
            zmap = q16(torch.rand(2, z, 4 * h, 4 * w, generator=g) * 2 - 1)
            x = torch.cat([zmap.contiguous().view(2, z * 16, h, w), lr], 1)
        else:
            x = lr
        out = wrapped(x)
        wt = torch.randn(out.shape, generator=g)
        (out * wt).sum().backward()
        for hk in hooks:
            hk.remove()
        assert min(margins) > 0.05, min(margins)
        grads = {'g:' + k: p.grad.numpy().astype(np.float32) for k, p in net.named_parameters()}
        save('wgrad_kinkfree_%s_train' % tag, x=x.numpy(), wt=wt.numpy(), out=out.detach().numpy(), cfg=np.array([32, 1, 4, z]),
             min_preact=np.array(min(margins)), **{'w:' + k: v for k, v in sd_np(net).items()}, **grads)


if __name__ == '__main__':
    main()
