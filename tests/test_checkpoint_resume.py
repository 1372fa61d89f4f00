"""Checkpoint / resume of the training model on CPU (no compute): `save` writes `<step>_G.pth` and `<step>_D.pth` in the
reference's format ({'model_state_dict', 'optimizer_state_dict'}); a model built with train.resume picks the latest pair up,
restores both Adam states and the step counter (models/SRRaGAN_model.py:732-771, base_model.py:114-131); without resume the
pretrained paths are used."""
import os

import torch


class ND(dict):
    def __missing__(self, k):
        return None


def _opt(tmp_path, **train_over):
    train = ND(pixel_weight=1.0, pixel_criterion='l1', gan_type='vanilla', gan_weight=5e-3, lr_G=1e-4, beta1_G=0.9, weight_decay_G=0,
               lr_D=2e-4, beta1_D=0.9, weight_decay_D=0, D_update_ratio=1, D_init_iters=0, lr_scheme='MultiStepLR', lr_steps=[1000], lr_gamma=0.5,
               grad_accumulation_steps_G=1, grad_accumulation_steps_D=1)
    train.update(train_over)
    return ND(model='srragan', scale=4, gpu_ids=None, is_train=True, range=[0, 1], train=train, datasets=ND(train=ND(patch_size=144, batch_size=2)),
              path=ND(models=str(tmp_path / 'models'), pretrained_model_G=None, log=str(tmp_path)),
              network_G=ND(which_model_G='RRDB_net', CEM_arch=1, latent_input=None, latent_input_domain=None, latent_channels=None,
                           norm_type=None, mode='CNA', nf=32, nb=1, in_nc=3, out_nc=3, gc=32, scale=4),
              network_D=ND(which_model_D='discriminator_vgg_128', norm_type='batch', act_type='leakyrelu', mode='CNA', nf=8, in_nc=3))


def _fake_adam_state(optimizer):
    for group in optimizer.param_groups:
        for p in group['params']:
            p.grad = torch.full_like(p, 0.01)
    optimizer.step()
    optimizer.zero_grad()


def test_save_and_resume_generator_and_discriminator(tmp_path):
    from models import create_model
    if torch.cuda.is_available():
        return      # the CPU suite's business; the GPU suite round-trips a trained critic in test_zz_discriminator_gpu.py
    torch.manual_seed(0)
    model = create_model(_opt(tmp_path), accumulation_steps_per_batch=2)
    assert model.D_exists and model.step == 0
    _fake_adam_state(model.optimizer_G)
    _fake_adam_state(model.optimizer_D)
    os.makedirs(str(tmp_path / 'models'), exist_ok=True)
    model.save(7)
    for name in ('7_G.pth', '7_D.pth'):
        ck = torch.load(os.path.join(str(tmp_path / 'models'), name), map_location='cpu')
        assert set(ck.keys()) == {'model_state_dict', 'optimizer_state_dict'} and len(ck['optimizer_state_dict']['state']) > 0
    # a fresh run without resume ignores the self-trained checkpoints ...
    torch.manual_seed(1)
    fresh = create_model(_opt(tmp_path), accumulation_steps_per_batch=2)
    assert fresh.step == 0 and len(fresh.optimizer_G.state) == 0
    gkey = 'generated_image_model.model.0.weight'
    assert not torch.equal(fresh.netG.state_dict()[gkey], model.netG.state_dict()[gkey])
    # ... a resumed one continues from them
    torch.manual_seed(2)
    resumed = create_model(_opt(tmp_path, resume=1), accumulation_steps_per_batch=2)
    assert resumed.step == (7 + 1) * 2
    for a, b in zip(resumed.netG.state_dict().items(), model.netG.state_dict().items()):
        assert a[0] == b[0] and torch.equal(a[1], b[1]), a[0]
    for a, b in zip(resumed.netD.state_dict().items(), model.netD.state_dict().items()):
        assert a[0] == b[0] and torch.equal(a[1], b[1]), a[0]
    for opt_new, opt_old in ((resumed.optimizer_G, model.optimizer_G), (resumed.optimizer_D, model.optimizer_D)):
        so, sn = opt_old.state_dict()['state'], opt_new.state_dict()['state']
        assert len(sn) == len(so) > 0
        assert all(torch.equal(sn[k]['exp_avg'], so[k]['exp_avg']) for k in so)
    assert resumed.optimizer_D.param_groups[0]['lr'] == 2e-4


def test_inference_model_picks_latest_self_trained_generator(tmp_path):
    from models import create_model
    if torch.cuda.is_available():
        return
    torch.manual_seed(0)
    model = create_model(_opt(tmp_path, gan_weight=None))
    os.makedirs(str(tmp_path / 'models'), exist_ok=True)
    model.save(3)
    with torch.no_grad():
        model.netG.generated_image_model.model[0].bias.add_(1.0)
    model.save(11)
    opt = _opt(tmp_path)
    opt['is_train'] = False
    test_model = create_model(opt)
    assert test_model.gradient_step_num == 11
    assert torch.equal(test_model.netG.state_dict()['generated_image_model.model.0.bias'], model.netG.state_dict()['generated_image_model.model.0.bias'])
    older = create_model(opt)
    older.load(max_step=5)
    assert older.gradient_step_num == 3


def test_resume_save_resume_round_trip_with_logs(tmp_path):
    """After a resume the scalars read back from logs.npz (lr_G, lr_D, D_verified, verified_D_saved) must be plain Python values:
    as 0-d numpy arrays they get pickled into the optimizers' param_groups and the next checkpoint cannot be loaded
    (reference casts with bool(), models/SRRaGAN_model.py:209)."""
    from models import create_model
    if torch.cuda.is_available():
        return
    torch.manual_seed(0)
    model = create_model(_opt(tmp_path), accumulation_steps_per_batch=1)
    _fake_adam_state(model.optimizer_G)
    _fake_adam_state(model.optimizer_D)
    os.makedirs(str(tmp_path / 'models'), exist_ok=True)
    model.log_dict['l_d_real'].append((2, 0.7))
    model.save(3)
    model.save_log()
    resumed = create_model(_opt(tmp_path, resume=1), accumulation_steps_per_batch=1)
    assert type(resumed.lr_G) is float and type(resumed.lr_D) is float
    assert type(resumed.D_verified) is bool and type(resumed.verified_D_saved) is bool
    assert all(type(g['lr']) is float for o in resumed.optimizers for g in o.param_groups)
    _fake_adam_state(resumed.optimizer_G)
    resumed.save(5)
    resumed.save_log()
    again = create_model(_opt(tmp_path, resume=1), accumulation_steps_per_batch=1)      # raised UnpicklingError before the cast
    assert again.step == 6 and len(again.optimizer_G.state) > 0
    again.load(max_step=4, resume_train=True)                                            # the lr roll-back path of update_learning_rate
    assert again.step == 4
