"""Latent-control loss (FilterLoss, structure-tensor descriptors) on CPU: the oracle against the reference's golden fixture, and
the mirror's host logic (normalisation, percentile history, autograd wiring) with the CUDA statistics kernel replaced by the
oracle's torch restatement."""
import pytest
import torch

from util import golden, rel_err

TAGS = ['SVDinNormedOut_structure_tensor', 'structure_tensor']


@pytest.mark.parametrize('tag', TAGS)
def test_oracle_filter_loss_matches_reference_golden(tag):
    from oracle import esr_oracle as O
    g = golden('filterloss_structure_tensor')
    history = [[], [], []]
    for call in range(2):
        sr = torch.from_numpy(g['%s:sr%d' % (tag, call)]).requires_grad_(True)
        out = O.filter_loss_structure_tensor(sr, torch.from_numpy(g['%s:hr%d' % (tag, call)]), torch.from_numpy(g['%s:z%d' % (tag, call)]),
                                             history, latent_channels=tag)
        assert rel_err(out.detach(), torch.from_numpy(g['%s:out%d' % (tag, call)]))[0] < 1e-5
    out.mean().backward()
    assert rel_err(sr.grad, torch.from_numpy(g['%s:gsr1' % tag]))[0] < 1e-4


@pytest.mark.parametrize('tag', TAGS)
def test_mirror_filter_loss_host_logic(monkeypatch, tag):
    from oracle import esr_oracle as O
    import models.modules.loss as loss
    monkeypatch.setattr(loss, 'structure_tensor_means', O.structure_tensor_means)
    g = golden('filterloss_structure_tensor')
    crit = loss.FilterLoss(latent_channels=tag)
    assert crit.num_channels == 3 and crit.built
    for call in range(2):
        sr = torch.from_numpy(g['%s:sr%d' % (tag, call)]).requires_grad_(True)
        out = crit({'SR': sr, 'HR': torch.from_numpy(g['%s:hr%d' % (tag, call)]), 'Z': torch.from_numpy(g['%s:z%d' % (tag, call)])})
        assert out.shape == (3, 3)
        assert rel_err(out.detach(), torch.from_numpy(g['%s:out%d' % (tag, call)]))[0] < 1e-5
    out.mean().backward()
    assert rel_err(sr.grad, torch.from_numpy(g['%s:gsr1' % tag]))[0] < 1e-4
    assert [len(h) for h in crit.collected_ratios] == [6, 6, 6]


def test_unbuilt_descriptors_fail_loudly():
    import models.modules.loss as loss
    crit = loss.FilterLoss(latent_channels='STD_directional')
    assert crit.num_channels == 3
    with pytest.raises(NotImplementedError):
        crit({'SR': torch.zeros(1, 3, 4, 4), 'HR': torch.zeros(1, 3, 4, 4), 'Z': torch.zeros(1, 3, 4, 4)})
    assert loss.FilterLoss(latent_channels=3).num_channels == 3 and loss.FilterLoss(latent_channels=None).num_channels == 0


def test_svd_2_latent_z_matches_definition():
    from models.SRRaGAN_model import SVD_2_LatentZ
    import numpy as np
    v = torch.tensor([[[[0.7]], [[0.2]], [[np.pi / 3]]]])
    z = SVD_2_LatentZ(v)
    s, c = np.sin(np.pi / 3), np.cos(np.pi / 3)
    ref = [2 * (0.2 * s * s + 0.7 * c * c) - 1, 2 * (0.7 * s * s + 0.2 * c * c) - 1, 2 * 0.5 * s * c]
    assert z.shape == (1, 3, 1, 1) and np.allclose(z.flatten().numpy(), ref, atol=1e-6)
