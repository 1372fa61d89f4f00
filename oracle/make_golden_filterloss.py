"""Latent-control loss golden fixture: the UNMODIFIED reference `FilterLoss` (models/modules/loss.py) in model-training mode
for both structure-tensor descriptors, called twice (the percentile history accumulates across calls), with the gradient of the
second call's mean with respect to the reconstructed image.  Build container only; the fixture is committed."""
import os
import sys


HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import torch  # noqa: E402
from models.modules.loss import FilterLoss  # noqa: E402
from make_golden import save  # noqa: E402


def main():
    g = torch.Generator().manual_seed(2024)
    arrays = {}
    for tag in ('SVDinNormedOut_structure_tensor', 'structure_tensor'):
        crit = FilterLoss(latent_channels=tag)
        outs = []
        for call in range(2):
            hr = torch.rand(3, 3, 20, 26, generator=g)
            sr = (hr + 0.2 * torch.randn(3, 3, 20, 26, generator=g)).requires_grad_(True)
            z = torch.rand(3, 3, 20, 26, generator=g) * 2 - 1
            out = crit({'SR': sr, 'HR': hr, 'Z': z})
            outs.append(out)
            arrays.update({'%s:sr%d' % (tag, call): sr.detach().numpy(), '%s:hr%d' % (tag, call): hr.numpy(), '%s:z%d' % (tag, call): z.numpy(),
                           '%s:out%d' % (tag, call): out.detach().numpy()})
        outs[1].mean().backward()
        arrays['%s:gsr1' % tag] = sr.grad.numpy()
    save('filterloss_structure_tensor', **arrays)


if __name__ == '__main__':
    main()
