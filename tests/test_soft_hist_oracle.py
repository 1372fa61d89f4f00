"""oracle/esr_oracle.soft_histogram* (the restatement the GPU kernels are held to) against the UNMODIFIED reference's SoftHistogramLoss
(oracle/make_golden_hist.py): loss values and image gradients of the grey-level `hist` and `dict` objectives."""
import numpy as np
import torch

from util import golden


def test_soft_histogram_oracle_matches_reference():
    from oracle import esr_oracle as O
    g = golden('soft_hist')
    cur = torch.from_numpy(g['cur']).requires_grad_(True)
    des, mask = torch.from_numpy(g['desired']), torch.from_numpy(g['image_mask'])
    for name, T, dic in (('hist', 5e-4, False), ('dict', 1e-3, True), ('hist_warm', 2e-2, False)):
        cur.grad = None
        loss = O.soft_histogram_loss_gray(cur, des, mask, T, dic)
        loss.mean().backward()
        assert np.allclose(loss.detach().numpy(), g[name + ':loss'], rtol=1e-5, atol=1e-9), name
        ref = torch.from_numpy(g[name + ':grad'])
        assert float((cur.grad - ref).abs().max()) < 1e-5 * float(ref.abs().max()), name
