"""The oracle (oracle/esr_oracle.py) against outputs of the unmodified reference (tests/golden, produced by
oracle/make_golden.py).  fp32 CPU on both sides: tolerance 1e-5 relative to the output range (observed ~1e-7;
the slack covers conv algorithm selection differences between builds of the same torch)."""
import numpy as np
import pytest
import torch

from oracle import esr_oracle as O
from util import golden, golden_state_dict, rel_err

TOL = 1e-5


@pytest.mark.parametrize('name,mode', [('rrdb_plain_x4', 'upconv'), ('rrdb_latent_x4', 'upconv'), ('rrdb_plain_x2', 'upconv'),
                                       ('rrdb_plain_x8', 'upconv'), ('rrdb_pixelshuffle_x4', 'pixelshuffle')])
def test_rrdbnet_matches_reference(name, mode):
    g = golden(name)
    nf, nb, s, z = [int(v) for v in g['cfg']]
    y = O.rrdbnet_forward(torch.from_numpy(g['x']), golden_state_dict(g), nf, nb, upscale=s, z=z, upsample_mode=mode)
    emax, el2 = rel_err(y, torch.from_numpy(g['y']))
    assert emax < TOL and el2 < TOL, (emax, el2)


@pytest.mark.parametrize('s', [2, 3, 4])
def test_cem_filters_and_projection_match_reference(s):
    g = golden('cem_x%d' % s)
    pre = {2: 0, 3: 1, 4: 1}[s]
    x, gi = torch.from_numpy(g['x_lr']), torch.from_numpy(g['g'])
    dk, ih = g['ds_kernel'], g['inv_hTh']
    assert rel_err(O.cem_down(gi, dk, s, pre), torch.from_numpy(g['down']))[0] < TOL
    assert rel_err(O.cem_inv(x, ih), torch.from_numpy(g['inv']))[0] < TOL
    assert rel_err(O.cem_up(x, dk, s, pre), torch.from_numpy(g['up']))[0] < TOL
    assert rel_err(O.cem_project(x, gi, dk, ih, s, pre), torch.from_numpy(g['out_train']))[0] < TOL
    # eval mode with a given generated image: pad both, project, crop (CEMnet.py:300-301,311)
    m_lr, m_hr = int(g['margins'][0]), int(g['margins'][1])
    xp = torch.nn.functional.pad(x, (m_lr,) * 4, mode='replicate')
    gp = torch.nn.functional.pad(gi, (m_hr,) * 4, mode='replicate')
    out = O.cem_project(xp, gp, dk, ih, s, pre)[:, :, m_hr:-m_hr, m_hr:-m_hr]
    assert rel_err(out, torch.from_numpy(g['out_eval']))[0] < TOL


@pytest.mark.parametrize('name,fixture', [('cem_rrdb_plain_x4', 'rrdb_plain_x4'), ('cem_rrdb_latent_x4', 'rrdb_latent_x4')])
def test_cem_wrapped_generator_matches_reference(name, fixture):
    g, gw, gc = golden(name), golden(fixture), golden('cem_x4')
    nf, nb, s, z = [int(v) for v in gw['cfg']]
    sd = golden_state_dict(gw, prefix='generated_image_model.')
    for mode, key in ((False, 'y_train'), (True, 'y_eval')):
        y = O.cem_wrapped_forward(torch.from_numpy(g['x']), sd, gc['ds_kernel'], gc['inv_hTh'], s, 1, int(gc['margins'][0]), mode,
                                  nf, nb, z=z)
        emax, el2 = rel_err(y, torch.from_numpy(g[key]))
        assert emax < TOL and el2 < TOL, (key, emax, el2)


def test_latent_packing_is_a_raw_view():
    z = torch.arange(2 * 3 * 8 * 12, dtype=torch.float32).view(2, 3, 8, 12)
    x = torch.zeros(2, 3, 2, 3)
    packed = O.pack_latent(z, x, 4)
    assert packed.shape == (2, 51, 2, 3)
    # bit-exact index permutation: undoing the view gives Z back
    assert torch.equal(packed[:, :48].contiguous().view(2, 3, 8, 12), z)


def test_cem_invariants_on_the_oracle():
    """Analytic known-answer checks: LR consistency and idempotence in the interior."""
    gc = golden('cem_x4')
    dk, ih = gc['ds_kernel'], gc['inv_hTh']
    gen = torch.Generator().manual_seed(7)
    x, gi = torch.rand(1, 3, 40, 40, generator=gen), torch.rand(1, 3, 160, 160, generator=gen)
    out = O.cem_project(x, gi, dk, ih, 4, 1)
    back = O.cem_down(out, dk, 4, 1)
    m = int(gc['margins'][0])
    assert (back - x)[:, :, m:-m, m:-m].abs().max() < 2e-5
    again = O.cem_project(x, out, dk, ih, 4, 1)
    assert (again - out)[:, :, 4 * m:-4 * m, 4 * m:-4 * m].abs().max() < 5e-5
    assert abs(float(dk.sum()) - 1) < 1e-6


@pytest.mark.parametrize('name,fixture,eval_mode', [('grad_cem_rrdb_latent_eval', 'rrdb_latent_x4', True),
                                                    ('grad_cem_rrdb_plain_train', 'rrdb_plain_x4', False),
                                                    ('grad_kinkfree_latent_eval', 'grad_kinkfree_latent_eval', True),
                                                    ('grad_kinkfree_latent_train', 'grad_kinkfree_latent_eval', False)])
def test_input_gradient_matches_reference_autograd(name, fixture, eval_mode):
    """d(sum(out*Wt))/dx through the oracle (torch autograd on the restatement) vs the reference's own autograd."""
    g, gw, gc = golden(name), golden(fixture), golden('cem_x4')
    nf, nb, s, z = [int(v) for v in gw['cfg']]
    sd = golden_state_dict(gw, prefix='generated_image_model.')
    x = torch.from_numpy(g['x']).clone().requires_grad_(True)
    out = O.cem_wrapped_forward(x, sd, gc['ds_kernel'], gc['inv_hTh'], s, 1, int(gc['margins'][0]), eval_mode, nf, nb, z=z)
    (out * torch.from_numpy(g['wt'])).sum().backward()
    emax, el2 = rel_err(x.grad, torch.from_numpy(g['gx']))
    assert emax < 1e-4 and el2 < 1e-4, (emax, el2)
