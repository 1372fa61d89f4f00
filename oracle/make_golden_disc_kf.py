"""Kink-free critic fixture: the UNMODIFIED reference `Discriminator_VGG_128` (CPU fp32, train-mode BatchNorm, base_nf 8) with every
LeakyReLU input kept away from zero, so that its gradients can be compared at ARITHMETIC precision (1e-3) by an implementation that
is not bit-identical.  (On the plain fixture of make_golden_disc.py pre-activations cross zero everywhere: every unit whose
pre-activation is smaller than the forward rounding error may legitimately take the other slope, and the gradient error of any two
implementations is ~sqrt(fraction of flipped units) - 1e-2 even for near-fp32 arithmetic - whatever the quality of their backward.)

Construction: conv0 (no norm) gets biases of magnitude 2..3 with random signs and small weights; every BatchNorm gets gamma 0.2 +- 10 %
and beta of magnitude 2..3 with random signs (the normalised activations are bounded by ~6 sigma, so |gamma * x_hat| < 1.5 < |beta|);
the hidden classifier layer gets biases 2..3 and small weights.  Both slopes stay exercised, channel by channel.
Build container only (`python oracle/make_golden_disc_kf.py`); the fixture is committed."""
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
import models.modules.architecture as arch  # noqa: E402
import models.networks as networks  # noqa: E402
from make_golden import save  # noqa: E402


def main():
    torch.manual_seed(78)
    g = torch.Generator().manual_seed(101)
    nf, n = 8, 4
    net = arch.Discriminator_VGG_128(in_nc=3, base_nf=nf, input_patch_size=128)
    with contextlib.redirect_stdout(io.StringIO()):
        networks.init_weights(net, 'kaiming', scale=1)
    pm = lambda shape: torch.where(torch.rand(shape, generator=g) < 0.5, -1.0, 1.0) * (2.0 + torch.rand(shape, generator=g))
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name == 'features.0.weight':
                p.mul_(0.3)
            elif name == 'features.0.bias':
                p.copy_(pm(p.shape))
            elif name.startswith('features') and p.dim() == 1 and name.endswith('weight'):     # BatchNorm gamma
                p.copy_(0.2 * (1 + 0.1 * (2 * torch.rand(p.shape, generator=g) - 1)))
            elif name.startswith('features') and name.endswith('bias'):
                idx = int(name.split('.')[1])
                if isinstance(net.features[idx], torch.nn.BatchNorm2d):
                    p.copy_(pm(p.shape))                                                          # BatchNorm beta
                else:
                    p.normal_(0, 0.1, generator=g)                                                # conv bias in front of a BatchNorm
            elif name == 'classifier.0.weight':
                p.mul_(0.1)
            elif name == 'classifier.0.bias':
                p.copy_(pm(p.shape))
            elif name == 'classifier.2.bias':
                p.normal_(0, 0.1, generator=g)
    net.train()
    margins = []
    hooks = [m.register_forward_pre_hook(lambda mod, inp: margins.append(float(inp[0].abs().min())))
             for m in net.modules() if isinstance(m, torch.nn.LeakyReLU)]
    w0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.rand(n, 3, 128, 128, generator=g).requires_grad_(True)
    out = net(x)
    wt = torch.randn(out.shape, generator=g)
    (out * wt).sum().backward()
    for h in hooks:
        h.remove()
    print('min |LeakyReLU input| = %.3f over %d activations' % (min(margins), len(margins)))
    assert min(margins) > 0.05
    arrays = {'w:' + k: v.numpy() for k, v in w0.items()}
    arrays.update({'g:' + k: p.grad.numpy().astype(np.float32) for k, p in net.named_parameters()})
    save('disc_vgg128_nf8_kf', x=x.detach().numpy(), wt=wt.numpy(), out=out.detach().numpy(), gx=x.grad.numpy(), cfg=np.array([nf, n]),
         min_preact=np.array(min(margins)), **arrays)


if __name__ == '__main__':
    main()
