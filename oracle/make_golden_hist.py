"""Soft-histogram fixture: the UNMODIFIED reference `SoftHistogramLoss` (Z_optimization.py:24-230) on CPU (its hard-coded
torch.device('cuda') / torch.cuda.*Tensor types redirected in memory, nothing edited) in the configurations that still run on the
installed torch: grey-level histogram (`hist`) and dictionary (`dict`) objectives with patch_size 1.  (The patch / KDE variants prune
their bins with `mask.any(1) ^ 1`, which modern torch turns into an integer index instead of a mask: they cannot run unmodified; the
kernels' patch mode is held to oracle/esr_oracle.soft_histogram, a restatement of :196-210 that these fixtures pin.)
Stores the loss and its gradient with respect to the image.  Build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import torch  # noqa: E402
from make_golden import save  # noqa: E402


def main():
    import Z_optimization as Zmod

    class TorchProxy:
        def __getattr__(self, k):
            return getattr(torch, k)

        def device(self, *a, **k):
            return torch.device('cpu')
    Zmod.torch = TorchProxy()
    g = torch.Generator().manual_seed(5)
    H, W = 24, 20
    desired = torch.rand(1, 3, H, W, generator=g) * 0.6 + 0.2
    cur = (torch.rand(2, 3, H, W, generator=g)).requires_grad_(True)
    desired_mask = (np.random.RandomState(1).rand(H, W) > 0.3)
    image_mask = torch.from_numpy((np.random.RandomState(2).rand(H, W) > 0.4).astype(np.float32))
    arrays = dict(desired=desired.numpy(), cur=cur.detach().numpy(), desired_mask=desired_mask, image_mask=image_mask.numpy())
    for name, kw in (('hist', dict(temperature=5e-4, dictionary_not_histogram=False)), ('dict', dict(temperature=1e-3, dictionary_not_histogram=True)),
                     ('hist_warm', dict(temperature=2e-2, dictionary_not_histogram=False))):
        loss_mod = Zmod.SoftHistogramLoss(bins=256, min=0, max=1, desired_hist_image=[desired], desired_hist_image_mask=[desired_mask],
                                          input_im_HR_mask=image_mask, gray_scale=True, patch_size=1, **kw)
        if cur.grad is not None:
            cur.grad = None
        loss = loss_mod(cur)
        loss.mean().backward()
        arrays[name + ':loss'] = loss.detach().numpy()
        arrays[name + ':grad'] = cur.grad.numpy().copy()
        print(name, loss.detach().numpy(), float(cur.grad.abs().max()))
    save('soft_hist', **arrays)


if __name__ == '__main__':
    main()
