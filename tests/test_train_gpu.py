"""Weight gradients of the CEM-wrapped generator (training backward) against the reference's own autograd.

Fixture: oracle/make_golden_wgrad.py (unmodified reference, CPU fp32, kink-free construction).  Tolerance: training runs
the engine on bf16 operands (activations, weights and gradients in ONE tensor-core format; fp16 gradients underflow on
the dense blocks' inner gradients, 1e-7 here) with fp32 accumulation and an fp32 trunk, like the reference's own
bf16-autocast training configuration (BASELINE config 3).  bf16 has an 8-bit mantissa and every one of the ~30 conv
stages of the forward and of the backward chain rounds its output to it: the forward output is held to 1e-2 of its
range, every gradient tensor to cosine >= 0.998, rel-L2 <= 6e-2 and max-abs <= 1e-1 of its range (measured worst
case is printed).  The kernels themselves are exact to 1e-5 on identical operands (tests/test_wgrad_gpu.py)."""
import pytest
import torch

from util import golden, mirror_rrdb, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(autouse=True)
def _watchdog():
    yield
    from esr_b200 import lib
    wd = lib.watchdog()
    assert wd[0] == 0, 'pipeline watchdog fired: %r' % (wd,)


@pytest.mark.parametrize('tag,poison', [('plain', False), ('latent', False), ('latent', True)])
def test_weight_gradients_match_reference_autograd(tag, poison, monkeypatch):
    """poison: the backward's scratch gradient buffers start as NaN instead of recycled memory (engine.backward allocates them without a
    zero-fill): a launch that read anything it had not written would poison the gradients"""
    from esr_b200 import ops
    if poison:
        monkeypatch.setenv('ESR_POISON', '1')
    from CEM.CEMnet import CEMnet, Get_CEM_Conf
    ops.device_check()
    g = golden('wgrad_kinkfree_%s_train' % tag)
    net = mirror_rrdb(g)
    wrapped = CEMnet(Get_CEM_Conf(4)).WrapArchitecture_PyTorch(net, None).to(DEV)
    wrapped.train()
    x = torch.from_numpy(g['x']).to(DEV)
    out = wrapped(x)
    ref_out = torch.from_numpy(g['out'])
    assert (out.detach().cpu() - ref_out).abs().max().item() < 1e-2 * max(1.0, ref_out.abs().max().item())
    (out * torch.from_numpy(g['wt']).to(DEV)).sum().backward()
    worst, bad = (0.0, 0.0, ''), []
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        ref = torch.from_numpy(g['g:' + name])
        emax, el2 = rel_err(p.grad.cpu(), ref)
        cos = torch.nn.functional.cosine_similarity(p.grad.cpu().flatten().double(), ref.flatten().double(), dim=0).item()
        worst = max(worst, (el2, emax, name))
        if not (el2 < 6e-2 and emax < 1e-1 and cos > 0.998):
            bad.append((name, round(emax, 5), round(el2, 5), round(cos, 5)))
    assert not bad, bad
    print('worst weight gradient: %s rel-L2 %.2e max %.2e' % (worst[2], worst[0], worst[1]))
    # the CEM filters stay frozen
    for name, p in wrapped.named_parameters():
        if 'Filter_OP' in name:
            assert p.grad is None


def test_sgd_step_reduces_l1_loss_and_accumulates():
    """two backward passes accumulate into .grad (gradient accumulation, SRRaGAN_model.py:392-396); one optimizer step
    along the gradient lowers the pixel loss"""
    from esr_b200 import ops
    import models.modules.architecture as arch
    ops.device_check()
    torch.manual_seed(3)
    net = arch.RRDBNet(3, 3, 32, 1, upscale=4, num_latent_channels=0).to(DEV)
    for p in net.parameters():
        torch.nn.init.normal_(p, 0, 0.03)
    x = torch.rand(2, 3, 24, 28, device=DEV)
    hr = torch.rand(2, 3, 96, 112, device=DEV)
    loss0 = (net(x) - hr).abs().mean()
    loss0.backward()
    g1 = [p.grad.clone() for p in net.parameters()]
    (net(x) - hr).abs().mean().backward()
    for a, p in zip(g1, net.parameters()):
        assert torch.allclose(p.grad, 2 * a, rtol=1e-4, atol=1e-7)
    opt = torch.optim.SGD(net.parameters(), lr=0.05)
    for p, a in zip(net.parameters(), g1):
        p.grad.copy_(a)
    opt.step()
    with torch.no_grad():
        loss1 = (net(x) - hr).abs().mean()
    assert loss1.item() < loss0.item(), (loss0.item(), loss1.item())


def _train_opt(tmp_path, **train_over):
    class ND(dict):
        def __missing__(self, k):
            return None
    train = ND(pixel_weight=1.0, pixel_criterion='l1', range_weight=0.1, lr_G=2e-4, beta1_G=0.9, weight_decay_G=0,
               lr_scheme='MultiStepLR', lr_steps=[1000], lr_gamma=0.5, grad_accumulation_steps_G=2, grad_accumulation_steps_D=2)
    train.update(train_over)
    return ND(model='srragan', scale=4, gpu_ids=[0], is_train=True, range=[0, 1], train=train,
              datasets=ND(train=ND(patch_size=128, batch_size=2)),
              path=ND(models=str(tmp_path / 'models'), pretrained_model_G=None, log=str(tmp_path)),
              network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input=None, latent_input_domain=None, latent_channels=None,
                           norm_type=None, mode='CNA', nf=32, nb=1, in_nc=3, out_nc=3, gc=32, scale=4),
              network_D=ND(which_model_D='discriminator_vgg_128', norm_type='batch', act_type='leakyrelu', mode='CNA', nf=16, in_nc=3))


def test_srragan_model_generator_training_step(tmp_path):
    """create_model(is_train) -> feed_data -> optimize_parameters, the calls train.py:69-116 makes: pixel + range loss,
    two accumulation steps per Adam step, first gradient step idle (as in the reference); the L1 loss goes down."""
    from esr_b200 import ops
    from models import create_model
    ops.device_check()
    torch.manual_seed(5)
    model = create_model(_train_opt(tmp_path), accumulation_steps_per_batch=2)
    assert model.netG.module.pre_pad is False
    lr = torch.rand(2, 3, 32, 32)
    hr = torch.nn.functional.interpolate(lr, scale_factor=4, mode='bicubic', align_corners=False).clamp(0, 1)
    w0 = [p.detach().clone() for p in model.netG.parameters() if p.requires_grad]
    for it in range(14):
        model.feed_data({'LR': lr, 'HR': hr})
        model.optimize_parameters()
        if it == 1:   # gradient step 0 is idle
            assert all(torch.equal(a, p.detach()) for a, p in zip(w0, [p for p in model.netG.parameters() if p.requires_grad]))
    log = model.log_dict['l_g_pix']
    assert len(log) == 6 and log[-1][1] < log[0][1], log
    assert model.fake_H.shape == (2, 3, 128 - 80, 128 - 80)        # HR_unpadder crops the invalid margins (SRRaGAN_model.py:333)
    assert any(not torch.equal(a, p.detach()) for a, p in zip(w0, [p for p in model.netG.parameters() if p.requires_grad]))
    model.update_learning_rate(3)
    assert model.get_current_learning_rate() == 2e-4


def test_srragan_model_refuses_unbuilt_losses(tmp_path):
    from models import create_model
    with pytest.raises(NotImplementedError):
        create_model(_train_opt(tmp_path, optimalZ_loss_weight=1.0, optimalZ_loss_type='hist'))


def test_srragan_model_perceptual_training_step(tmp_path):
    """pixel + VGG-feature loss (SRRaGAN_model.py:434-451): real features detached, fake features carry the gradient
    through the frozen extractor into the generator"""
    from esr_b200 import ops
    from models import create_model
    ops.device_check()
    torch.manual_seed(6)
    model = create_model(_train_opt(tmp_path, feature_weight=1.0, feature_criterion='l1', grad_accumulation_steps_G=1, range_weight=None,
                                    lr_G=1e-3), accumulation_steps_per_batch=1)
    assert all(not p.requires_grad for p in model.netF.parameters())
    lr = torch.rand(2, 3, 36, 36)                      # 144 - 80 = 64: divisible by 16 for the four poolings
    hr = torch.nn.functional.interpolate(lr, scale_factor=4, mode='bicubic', align_corners=False).clamp(0, 1)
    for it in range(16):
        model.feed_data({'LR': lr, 'HR': hr})
        model.optimize_parameters()
    fea, pix = model.log_dict['l_g_fea'], model.log_dict['l_g_pix']
    assert len(fea) == 15 and all(torch.isfinite(torch.tensor(v)) for _, v in fea)
    # fifteen Adam steps (lr 1e-3) on a random extractor: the weighted objective goes down (single steps may wobble: bf16)
    total = [fea[k][1] + 1.0 * pix[k][1] for k in range(15)]
    assert sum(total[-3:]) / 3 < total[0], total
