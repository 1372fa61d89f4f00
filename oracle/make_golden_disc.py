"""Discriminator golden fixture: runs the UNMODIFIED reference `Discriminator_VGG_128` (CPU fp32, train-mode BatchNorm) on a
seeded batch and stores logits, d(loss)/d(image), d(loss)/d(parameter) for loss = sum(logits * wt), the running statistics after
that forward, and the eval-mode logits computed with them.  Build container only (`python oracle/make_golden_disc.py`); the
fixture is committed.  base_nf = 8 keeps it small (316 k parameters); all operands are fp16-exact."""
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
from make_golden import q16, save  # noqa: E402


def main():
    torch.manual_seed(77)
    g = torch.Generator().manual_seed(99)
    nf, n = 8, 4
    net = arch.Discriminator_VGG_128(in_nc=3, base_nf=nf, input_patch_size=128)
    with contextlib.redirect_stdout(io.StringIO()):
        networks.init_weights(net, 'kaiming', scale=1)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith('bias'):
                p.normal_(0, 0.1, generator=g)
            elif p.dim() == 1:          # BatchNorm weight
                p.copy_(1 + 0.2 * torch.randn(p.shape, generator=g))
            p.copy_(q16(p))
    net.train()
    w0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = q16(torch.rand(n, 3, 128, 128, generator=g)).requires_grad_(True)
    out = net(x)
    wt = torch.randn(out.shape, generator=g)
    (out * wt).sum().backward()
    after = {k: v.detach().clone() for k, v in net.state_dict().items() if 'running' in k}
    net.eval()
    with torch.no_grad():
        out_eval = net(x)
    arrays = {'w:' + k: (v.numpy().astype(np.float16) if v.dtype.is_floating_point and 'running' not in k else v.numpy()) for k, v in w0.items()}
    arrays.update({'r:' + k: v.numpy() for k, v in after.items()})
    arrays.update({'g:' + k: p.grad.numpy().astype(np.float32) for k, p in net.named_parameters()})
    save('disc_vgg128_nf8', x=x.detach().numpy().astype(np.float16), wt=wt.numpy(), out=out.detach().numpy(), out_eval=out_eval.numpy(), gx=x.grad.numpy(),
         cfg=np.array([nf, n]), **arrays)


if __name__ == '__main__':
    main()
