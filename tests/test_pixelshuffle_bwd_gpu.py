"""Backward through the pixel-shuffle up-sampler (models/modules/block.py:278-291: conv nf->4nf, nn.PixelShuffle(2), LeakyReLU): the
un-shuffle kernel is a bit-exact permutation, and the generator's input / weight gradients in 'pixelshuffle' mode match autograd of the
oracle's restatement (fp64, CPU) in the parity mode at 1e-3 on a kink-free construction (every LeakyReLU input away from zero)."""
import re

import pytest
import torch
import torch.nn.functional as F

from util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def test_pixel_unshuffle_is_the_adjoint_permutation():
    from esr_b200 import ops
    ops.device_check()
    torch.manual_seed(0)
    x = torch.randn(2, 16, 12, 20, device=DEV)                      # [N, 4c, h, w]
    shuffled = F.pixel_shuffle(x, 2)                                # [N, c, 2h, 2w]
    s16, _ = ops.pack_nchw(shuffled, dtype=torch.float16)
    back = ops.unpack_planes(ops.pixel_unshuffle2(s16), 16)
    assert torch.equal(back, x.half().float())
    # split tensors: both halves are permuted
    s2, _ = ops.pack_nchw(shuffled, dtype=ops.SPLIT)
    back2 = ops.unpack_planes(ops.pixel_unshuffle2(s2, dtype=ops.SPLIT), 16, split=True)
    assert (back2 - x).abs().max().item() < 1e-4 * x.abs().max().item()


@pytest.mark.parametrize('scale', [2, 4])
def test_pixelshuffle_generator_gradients_parity_mode(scale):
    import models.modules.architecture as arch
    from esr_b200 import ops, precision
    from oracle import esr_oracle as O
    ops.device_check()
    torch.manual_seed(3 + scale)
    nf, nb = 32, 1
    net = arch.RRDBNet(3, 3, nf, nb, upscale=scale, upsample_mode='pixelshuffle', num_latent_channels=0)
    g = torch.Generator().manual_seed(11)
    n_up = {2: 1, 4: 2}[scale]
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith('weight'):
                p.mul_(0.5)
            else:
                # activated convs (growth convs, shuffle convs, HR_conv0): biases of magnitude 2..3 keep the LeakyReLU inputs off the kink
                activated = bool(re.search(r'convs\.[0-3]\.0\.bias$', name) or re.match(r'model\.[2-9]\.0\.bias$', name) or name == 'model.%d.bias' % (2 + n_up))
                if activated:
                    sign = torch.where(torch.rand(p.shape, generator=g) < 0.5, -1.0, 1.0)
                    p.copy_(sign * (2.0 + torch.rand(p.shape, generator=g)))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))
    x = torch.rand(2, 3, 24, 20, generator=g)
    wt = torch.randn(2, 3, 24 * scale, 20 * scale, generator=g)
    # oracle (fp64 autograd on the CPU)
    sd = {k: v.detach().double().requires_grad_(True) for k, v in net.state_dict().items()}
    xr = x.double().requires_grad_(True)
    ref = O.rrdbnet_forward(xr, sd, nf, nb, upscale=scale, z=0, upsample_mode='pixelshuffle')
    (ref * wt.double()).sum().backward()
    # CUDA path, parity mode
    net = net.to(DEV)
    for p in net.parameters():
        p.requires_grad_(True)
    xd = x.to(DEV).requires_grad_(True)
    with precision.use('parity'):
        out = net(xd)
        (out * wt.to(DEV)).sum().backward()
    emax, el2 = rel_err(out.detach().cpu(), ref.detach().float())
    assert emax < 1e-3 and el2 < 1e-3, (emax, el2)
    emax, el2 = rel_err(xd.grad.cpu(), xr.grad.float())
    print('pixelshuffle x%d: input gradient max %.2e rel-L2 %.2e' % (scale, emax, el2))
    assert emax < 1e-3 and el2 < 1e-3
    worst = 0.0
    scale_w = max(sd[k].grad.abs().max().item() for k in sd if k.endswith('weight'))
    for name, p in net.named_parameters():
        r = sd[name].grad.float()
        assert p.grad is not None, name
        if name.endswith('bias'):      # biases: relative to the weight-gradient scale when the reference is tiny
            err = (p.grad.cpu() - r).abs().max().item() / max(r.abs().max().item(), 1e-3 * scale_w)
        else:
            err = rel_err(p.grad.cpu(), r)[1]
        worst = max(worst, err)
        assert err < 1e-3, (name, err)
    print('pixelshuffle x%d: worst parameter gradient %.2e' % (scale, worst))
