"""SRRaGANModel (inference surface) and Z_optimizer on the GPU against an oracle run of the same loop.

The reference's Z_optimizer hard-codes torch.device('cuda') (Z_optimization.py:32,369) so it cannot produce golden
vectors in the CPU-only build container; the oracle restates its loop (oracle/esr_oracle.z_optimize_l1) on top of the
golden-pinned forward."""
import os

import numpy as np
import pytest
import torch

from util import golden, golden_state_dict, rel_err

pytestmark = pytest.mark.gpu


def _opt(tmp_path, ckpt):
    from collections import defaultdict

    class ND(dict):
        def __missing__(self, k):
            return None
    return ND(model='srragan', scale=4, gpu_ids=[0], is_train=False, range=[0, 1],
              path=ND(models=str(tmp_path / 'models'), pretrained_model_G=ckpt, log=str(tmp_path)),
              network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input='all_layers', latent_input_domain='HR_downscaled',
                           latent_channels=3, norm_type=None, mode='CNA', nf=32, nb=1, in_nc=3, out_nc=3, gc=32, scale=4))


def _model(tmp_path):
    from models import create_model
    g = golden('rrdb_latent_x4')
    # a checkpoint in the reference's file format, of the BARE generator (exercises Adjust_State_Dict_Keys)
    ckpt = str(tmp_path / 'G.pth')
    torch.save({'model_state_dict': golden_state_dict(g), 'optimizer_state_dict': {}}, ckpt)
    return create_model(_opt(tmp_path, ckpt)), g


def test_model_surface_and_checkpoint_loading(tmp_path):
    from esr_b200 import ops
    ops.device_check()
    from oracle import esr_oracle as O
    model, g = _model(tmp_path)
    assert type(model.netG).__name__ == 'SingleDeviceDataParallel' and type(model.netG.module).__name__ == 'CEM_PyTorch'
    assert hasattr(model.netG.module, 'DownscaleOP') and model.num_latent_channels == 3 and model.Z_size_factor == 4
    gc, ge = golden('cem_x4'), golden('cem_rrdb_latent_x4')
    x = torch.from_numpy(ge['x'])
    z_hr = x[:, :48].contiguous().view(1, 3, 64, 48)
    model.feed_data({'LR': x[:, 48:], 'Z': z_hr}, need_GT=False)
    assert torch.equal(model.GetLatent().cpu(), z_hr)              # raw-view round trip, bit exact
    assert torch.equal(model.model_input.cpu(), x)
    model.test()
    assert (model.fake_H.cpu() - torch.from_numpy(ge['y_eval'])).abs().max().item() < 1e-3
    assert model.netG.module.pre_pad is False                      # test() restores train mode like the reference
    # scalar Z broadcasts over the HR grid (SRRaGAN_model.py:268-271)
    model.feed_data({'LR': x[:, 48:], 'Z': 0}, need_GT=False)
    assert model.GetLatent().shape == (1, 3, 64, 48) and float(model.GetLatent().abs().max()) == 0.0
    # save / load round trip in the reference's checkpoint format
    os.makedirs(model.save_dir, exist_ok=True)
    path = model.save('7')
    blob = torch.load(path)
    assert set(blob.keys()) == {'model_state_dict', 'optimizer_state_dict'}
    assert list(blob['model_state_dict'].keys())[0] == 'generated_image_model.model.0.weight'
    with pytest.raises(NotImplementedError):
        model.optimize_parameters()


@pytest.mark.parametrize('batch', [1, 2])
def test_z_optimizer_l1_matches_oracle_loop(tmp_path, batch):
    from esr_b200 import ops
    ops.device_check()
    from oracle import esr_oracle as O
    from Z_optimization import Z_optimizer
    model, g = _model(tmp_path)
    gc = golden('cem_x4')
    gen = torch.Generator().manual_seed(5)
    x_lr = torch.rand(1, 3, 16, 12, generator=gen)
    desired = torch.rand(1, 3, 64, 48, generator=gen)
    iters, lr, Z_range = 6, 0.1, 1.0
    sd = golden_state_dict(g, prefix='generated_image_model.')
    ref_losses, ref_Z = O.z_optimize_l1(sd, gc['ds_kernel'], gc['inv_hTh'], 4, 1, int(gc['margins'][0]), 32, 1, 3, x_lr, desired, Z_range, lr,
                                        iters, batch=batch)
    data = {'LR': x_lr.cuda().expand(batch, -1, -1, -1).contiguous(), 'desired': desired.cuda()}
    model.feed_data({'LR': data['LR'], 'Z': 0}, need_GT=False)
    model.test()
    zo = Z_optimizer(objective='l1', Z_size=[64, 48], model=model, Z_range=Z_range, max_iters=iters, data=data, initial_LR=lr, batch_size=batch)
    Z = zo.optimize()
    print('losses mine', ['%.5f' % v for v in zo.loss_values])
    print('losses ref ', ['%.5f' % v for v in ref_losses])
    assert len(zo.loss_values) == len(ref_losses)
    assert np.allclose(zo.loss_values, ref_losses, rtol=2e-3, atol=1e-5)
    assert zo.loss_values[-1] < zo.loss_values[0]                  # it actually optimises
    # Adam's normalised step turns every near-zero gradient component (|out-desired| and clamp kinks) into a +-lr move,
    # so individual latent pixels may legitimately differ; the bulk of the map must agree
    # (measured: ~10 % of the pixels, those whose gradient is below ~1e-3 of the largest one, end up > 0.02 apart while
    # the loss curves agree to 5 digits)
    close = ((Z.cpu() - ref_Z).abs() < 0.02).float().mean().item()
    cos = torch.nn.functional.cosine_similarity(Z.cpu().flatten(), ref_Z.expand_as(Z.cpu()).flatten(), dim=0).item()
    print('fraction of Z within 0.02 of the oracle: %.4f, cosine %.4f' % (close, cos))
    assert close > 0.8 and cos > 0.95, (close, cos)
    # generator parameters get their requires_grad status back, and no weight gradient was produced
    assert all(p.grad is None for p in model.netG.parameters())


def test_z_optimizer_other_objectives_run(tmp_path):
    from esr_b200 import ops
    ops.device_check()
    from Z_optimization import Z_optimizer
    model, g = _model(tmp_path)
    x_lr = torch.rand(1, 3, 16, 12, generator=torch.Generator().manual_seed(6)).cuda()
    for objective, kw in (('max_STD', {}), ('TV', {}), ('STD_increase', {'STD_increment': 0.01}), ('random_l1', {}),
                          ('periodicity', {'periodicity_points': [[3, 2]]}), ('nonInt_periodicity_Plus', {'periodicity_points': [[2.5, 3.3]], 'STD_increment': 0.01})):
        bs = 3 if 'random' in objective else 1
        data = {'LR': x_lr.expand(bs, -1, -1, -1).contiguous(), **kw}
        model.feed_data({'LR': data['LR'], 'Z': 0}, need_GT=False)
        model.test()
        zo = Z_optimizer(objective=objective, Z_size=[64, 48], model=model, Z_range=1.0, max_iters=4, data=data, initial_LR=0.1, batch_size=bs,
                         random_Z_inits='random' in objective)
        Z = zo.optimize()
        assert Z.shape == (bs, 3, 64, 48) and torch.isfinite(Z).all() and len(zo.loss_values) >= 1
    # the scribble tool on the device: region masks, labels 1 (colour) / 2 / 3 (brighten / darken) / 4 (smoothing), the loss must go down
    import numpy as np
    image_mask = np.zeros((64, 48), dtype=np.float32)
    image_mask[8:56, 6:42] = 1
    labels = np.zeros((64, 48), dtype=np.int64)
    labels[10:20, 8:30], labels[24:30, 8:20], labels[24:30, 24:38], labels[36:50, 10:36] = 1, 2, 3, 4
    data = {'LR': x_lr, 'desired': torch.rand(1, 3, 64, 48, generator=torch.Generator().manual_seed(8)).cuda(), 'scribble_mask': labels,
            'brightness_factor': 0.3}
    model.feed_data({'LR': x_lr, 'Z': 0}, need_GT=False)
    model.test()
    zo = Z_optimizer(objective='scribble', Z_size=[64, 48], model=model, Z_range=1.0, max_iters=8, data=data, initial_LR=0.1, batch_size=1,
                     image_mask=image_mask, Z_mask=1 * image_mask, initial_Z=1 * model.GetLatent())
    Z = zo.optimize()
    assert torch.isfinite(Z).all() and zo.loss_values[-1] <= zo.loss_values[0]
    with pytest.raises(NotImplementedError):
        Z_optimizer(objective='desired_SVD', Z_size=[64, 48], model=model, Z_range=1.0, max_iters=4, data={}, initial_LR=0.1)


def test_z_optimizer_graph_replay_matches_eager(tmp_path, monkeypatch):
    """the CUDA-graph replay of the iteration (launch-bound at GUI region sizes) is the same computation as the eager loop: same
    loss curve and latent map up to the rounding of Adam's capturable (tensor-valued step) arithmetic"""
    from esr_b200 import ops
    ops.device_check()
    from Z_optimization import Z_optimizer
    model, g = _model(tmp_path)
    gen = torch.Generator().manual_seed(9)
    x_lr = torch.rand(2, 3, 16, 12, generator=gen).cuda()
    desired = torch.rand(1, 3, 64, 48, generator=gen).cuda()
    runs = []
    for flag in ('1', '0'):
        monkeypatch.setenv('ESR_ZOPT_GRAPH', flag)
        data = {'LR': x_lr, 'desired': desired}
        model.feed_data({'LR': x_lr, 'Z': 0}, need_GT=False)
        model.test()
        zo = Z_optimizer(objective='l1', Z_size=[64, 48], model=model, Z_range=1.0, max_iters=12, data=data, initial_LR=0.1, batch_size=2)
        Z = zo.optimize()
        runs.append((list(zo.loss_values), Z.clone(), zo._graph_ok))
    assert runs[0][2] is True and runs[1][2] is False          # the first run really replayed a graph
    assert len(runs[0][0]) == len(runs[1][0])
    assert np.allclose(runs[0][0], runs[1][0], rtol=1e-3, atol=1e-6), (runs[0][0], runs[1][0])
    assert runs[0][0][-1] < runs[0][0][3] < runs[0][0][0]          # the replayed iterations keep optimising
    # Adam turns every near-zero gradient component into a +-lr move, so single latent pixels may differ (see the oracle test)
    close = ((runs[0][1] - runs[1][1]).abs() < 0.02).float().mean().item()
    cos = torch.nn.functional.cosine_similarity(runs[0][1].flatten(), runs[1][1].flatten(), dim=0).item()
    assert close > 0.8 and cos > 0.95, (close, cos)
