"""Latent-control loss on the B200: the structure-tensor kernels (forward + backward) against the oracle, FilterLoss against the
reference's golden fixture, and a latent-input training step of SRRaGANModel with latent_weight set."""
import numpy as np
import pytest
import torch

from util import golden, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('n,c,h,w', [(3, 3, 20, 26), (1, 1, 2, 2), (4, 3, 128, 128), (2, 5, 33, 7)])
def test_structure_tensor_kernels_match_oracle(n, c, h, w):
    from esr_b200 import ops
    from oracle import esr_oracle as O
    ops.device_check()
    g = torch.Generator().manual_seed(h * w)
    x = torch.rand(n, c, h, w, generator=g)
    xr = x.double().requires_grad_(True)
    ref = O.structure_tensor_means(xr)
    own = ops.structure_tensor(x.to(DEV))
    assert rel_err(own.cpu(), ref.detach())[0] < 1e-5
    gy = torch.randn(n, 3, generator=g)
    (ref * gy.double()).sum().backward()
    gx = ops.structure_tensor_bwd(x.to(DEV), gy.to(DEV))
    assert rel_err(gx.cpu(), xr.grad)[0] < 1e-5


@pytest.mark.parametrize('tag', ['SVDinNormedOut_structure_tensor', 'structure_tensor'])
def test_filter_loss_matches_reference_golden(tag):
    import models.modules.loss as loss
    g = golden('filterloss_structure_tensor')
    crit = loss.FilterLoss(latent_channels=tag)
    for call in range(2):
        sr = torch.from_numpy(g['%s:sr%d' % (tag, call)]).to(DEV).requires_grad_(True)
        out = crit({'SR': sr, 'HR': torch.from_numpy(g['%s:hr%d' % (tag, call)]).to(DEV), 'Z': torch.from_numpy(g['%s:z%d' % (tag, call)]).to(DEV)})
        assert rel_err(out.detach().cpu(), torch.from_numpy(g['%s:out%d' % (tag, call)]))[0] < 1e-4
    out.mean().backward()
    assert rel_err(sr.grad.cpu(), torch.from_numpy(g['%s:gsr1' % tag]))[0] < 1e-4


def test_srragan_model_latent_control_training_step(tmp_path):
    """explorable-SR configuration (options/train/train_explorable_SR.json): latent input in every layer, spatially uniform Z
    sampled per image, pixel + latent-control loss; the loss is finite, reaches the generator's latent-input weights and the
    optimiser moves them."""
    from esr_b200 import ops
    from models import create_model
    ops.device_check()

    class ND(dict):
        def __missing__(self, k):
            return None
    train = ND(pixel_weight=1.0, pixel_criterion='l1', latent_weight=1.0, lr_G=5e-4, beta1_G=0.9, weight_decay_G=0, lr_scheme='MultiStepLR',
               lr_steps=[1000], lr_gamma=0.5, grad_accumulation_steps_G=1, grad_accumulation_steps_D=1)
    opt = ND(model='srragan', scale=4, gpu_ids=[0], is_train=True, range=[0, 1], train=train, datasets=ND(train=ND(patch_size=128, batch_size=4)),
             path=ND(models=str(tmp_path / 'models'), pretrained_model_G=None, log=str(tmp_path)),
             network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input='all_layers', latent_input_domain='HR_downscaled',
                          latent_channels='SVDinNormedOut_structure_tensor', norm_type=None, mode='CNA', nf=32, nb=1, in_nc=3, out_nc=3, gc=32, scale=4))
    torch.manual_seed(9)
    model = create_model(opt)
    assert model.num_latent_channels == 3 and model.cri_latent is not None
    lr = torch.rand(4, 3, 32, 32)
    hr = torch.nn.functional.interpolate(lr, scale_factor=4, mode='bicubic', align_corners=False).clamp(0, 1)
    w0 = model.netG.module.generated_image_model.model[0].weight.detach().clone()
    for it in range(25):
        model.feed_data({'LR': lr, 'HR': hr})
        assert model.model_input.shape == (4, 3 * 16 + 3, 32, 32)
        model.optimize_parameters()
    w1 = model.netG.module.generated_image_model.model[0].weight.detach()
    assert not torch.equal(w0[:, :3], w1[:, :3])          # the latent channels are the first input channels of every conv
    logs = [np.array([v for _, v in model.log_dict['l_g_latent_%d' % ch]]) for ch in range(3)]
    assert all(len(l) == 24 and np.isfinite(l).all() for l in logs), logs
    total = sum(logs)
    print('latent-control loss, first / last five steps: %.4f / %.4f' % (total[:5].mean(), total[-5:].mean()))
