"""WGAN-GP fixture: the UNMODIFIED reference's GradientPenaltyLoss (models/modules/loss.py:260-279) on its own Discriminator_VGG_128
(train-mode BatchNorm, base_nf 8) with the kink-free weights of tests/golden/disc_vgg128_nf8_kf.npz: the penalty value and its
gradient with respect to every critic parameter (the double backward through the critic, models/SRRaGAN_model.py:362-371), plus the
gradient of (logits * wt).sum() + 10 * penalty, the way the D step combines them.  Build container only; the fixture is committed."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import torch  # noqa: E402
import models.modules.architecture as arch  # noqa: E402
from models.modules.loss import GradientPenaltyLoss  # noqa: E402
from make_golden import save  # noqa: E402


def main():
    src = np.load(os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'disc_vgg128_nf8_kf.npz'))
    nf, n = [int(v) for v in src['cfg']]
    net = arch.Discriminator_VGG_128(in_nc=3, base_nf=nf, input_patch_size=128)
    sd = {k[2:]: torch.from_numpy(src[k]) for k in src.files if k.startswith('w:')}
    net.load_state_dict(sd, strict=True)
    net.train()
    g = torch.Generator().manual_seed(7)
    real, fake = torch.rand(n, 3, 128, 128, generator=g), torch.rand(n, 3, 128, 128, generator=g)
    alpha = torch.rand(n, 1, 1, 1, generator=g)
    interp = (alpha * fake + (1 - alpha) * real).requires_grad_(True)       # SRRaGAN_model.py:364-366
    cri_gp = GradientPenaltyLoss(device=torch.device('cpu'))
    crit = net(interp)
    l_gp = cri_gp(interp, crit)
    grads = torch.autograd.grad(l_gp, list(net.parameters()), allow_unused=True, retain_graph=True)
    arrays = {'g:' + k: (gr.numpy() if gr is not None else np.zeros(tuple(p.shape), np.float32)) for (k, p), gr in zip(net.named_parameters(), grads)}
    # the input gradient whose norm is penalised, for the first-order check
    gx = torch.autograd.grad(crit, interp, torch.ones_like(crit), retain_graph=True)[0]
    save('wgan_gp_nf8_kf', interp=interp.detach().numpy(), l_gp=np.array(float(l_gp)), gx=gx.numpy(),
         norms=gx.view(n, -1).norm(2, dim=1).numpy(), cfg=np.array([nf, n]), **arrays)


if __name__ == '__main__':
    main()
