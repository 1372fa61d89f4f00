"""Network factory with the reference's entry points (models/networks.py:85-202): define_G / define_D /
define_F, kaiming initialisation (x0.1 for G), CEM wrapping and the DataParallel-shaped return value."""
import functools

import torch
import torch.nn as nn
from torch.nn import init

import models.modules.architecture as arch


def weights_init_normal(m, std=0.02):
    name = m.__class__.__name__
    if name.find('Conv') != -1 or name.find('Linear') != -1:
        init.normal_(m.weight.data, 0.0, std)
        if m.bias is not None:
            m.bias.data.zero_()
    elif name.find('BatchNorm2d') != -1:
        init.normal_(m.weight.data, 1.0, std)
        init.constant_(m.bias.data, 0.0)


def weights_init_kaiming(m, scale=1):
    # the CEM's fixed filters are tagged and must keep their designed taps (networks.py:30-31)
    if 'filter_layer' in m.__dict__ and m.__getattribute__('filter_layer'):
        return
    name = m.__class__.__name__
    if name.find('Conv') != -1 or name.find('Linear') != -1:
        init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
        m.weight.data *= scale
        if m.bias is not None:
            m.bias.data.zero_()
    elif 'BatchNorm2d' in name:
        init.constant_(m.weight.data, 1.0)
        init.constant_(m.bias.data, 0.0)


def weights_init_orthogonal(m):
    name = m.__class__.__name__
    if name.find('Conv') != -1 or name.find('Linear') != -1:
        init.orthogonal_(m.weight.data, gain=1)
        if m.bias is not None:
            m.bias.data.zero_()
    elif name.find('BatchNorm2d') != -1:
        init.constant_(m.weight.data, 1.0)
        init.constant_(m.bias.data, 0.0)


def init_weights(net, init_type='kaiming', scale=1, std=0.02):
    print('initialization method [{:s}]'.format(init_type))
    if init_type == 'normal':
        net.apply(functools.partial(weights_init_normal, std=std))
    elif init_type == 'kaiming':
        net.apply(functools.partial(weights_init_kaiming, scale=scale))
    elif init_type == 'orthogonal':
        net.apply(weights_init_orthogonal)
    else:
        raise NotImplementedError('initialization method [{:s}] not implemented'.format(init_type))


class SingleDeviceDataParallel(nn.DataParallel):
    """What the reference returns is `nn.DataParallel(netG)` and callers reach into `.module`
    (GUI.py:1687,2516).  esr_b200 runs one process per GPU (torch.distributed + NCCL), so the wrapper keeps
    the DataParallel type and `.module` attribute but always runs the wrapped module on its own device —
    no replicate / scatter / gather."""

    def __init__(self, module):
        super(SingleDeviceDataParallel, self).__init__(module, device_ids=[torch.cuda.current_device()])

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)


def define_G(opt, CEM=None, num_latent_channels=None, **kwargs):
    gpu_ids = opt['gpu_ids']
    opt_net = opt['network_G']
    which_model = opt_net['which_model_G']
    opt_net['latent_input'] = opt_net['latent_input'] if opt_net['latent_input'] != "None" else None
    if which_model == 'RRDB_net':
        latent = (opt_net['latent_input'] + '_' + opt_net['latent_input_domain']) if opt_net['latent_input'] is not None else None
        netG = arch.RRDBNet(in_nc=opt_net['in_nc'], out_nc=opt_net['out_nc'], nf=opt_net['nf'], nb=opt_net['nb'], gc=opt_net['gc'],
                            upscale=opt_net['scale'], norm_type=opt_net['norm_type'], act_type='leakyrelu', mode=opt_net['mode'],
                            upsample_mode='upconv', latent_input=latent, num_latent_channels=num_latent_channels)
    else:
        raise NotImplementedError('Generator model [{:s}] not recognized (esr_b200 builds RRDB_net only)'.format(which_model))
    if opt_net['CEM_arch']:
        netG = CEM.WrapArchitecture_PyTorch(netG, opt['datasets']['train']['patch_size'] if opt['is_train'] else None)
    if opt['is_train']:
        init_weights(netG, init_type='kaiming', scale=0.1)
    if torch.cuda.is_available():
        netG = netG.cuda()
    if gpu_ids:
        assert torch.cuda.is_available()
        netG = SingleDeviceDataParallel(netG)
    return netG


def define_D(opt, CEM=None, **kwargs):
    """models/networks.py:125-182: the critic sees HR patches with the CEM's invalid margins cropped; kaiming init (scale 1)"""
    gpu_ids = opt['gpu_ids']
    opt_net = opt['network_D']
    which_model = opt_net['which_model_D']
    input_patch_size = opt['datasets']['train']['patch_size']
    in_nc = opt_net['in_nc']
    assert not ((opt_net['pre_clipping'] or opt_net['decomposed_input']) and which_model != 'PatchGAN'), 'Unsupported yet'
    if CEM is not None:
        input_patch_size -= 2 * CEM.invalidity_margins_HR
    if which_model == 'discriminator_vgg_128':
        kw = {}
        if 'num_2_strides' in opt_net and opt_net['num_2_strides'] is not None:
            kw['num_2_strides'] = opt_net['num_2_strides']
        netD = arch.Discriminator_VGG_128(in_nc=in_nc, base_nf=opt_net['nf'], nb=opt_net['n_layers'], norm_type=opt_net['norm_type'],
                                          mode=opt_net['mode'], act_type=opt_net['act_type'], input_patch_size=input_patch_size, **kw)
    else:
        raise NotImplementedError('Discriminator model [{:s}] not recognized (esr_b200 builds discriminator_vgg_128 only)'.format(which_model))
    init_weights(netD, init_type='kaiming', scale=1)
    if torch.cuda.is_available():
        netD = netD.cuda()
    if gpu_ids:
        netD = SingleDeviceDataParallel(netD)
    return netD


def define_F(opt, use_bn=False, **kwargs):
    """models/networks.py:185-202: VGG19-54 before ReLU (feature_layer 34), input normalisation on, eval mode"""
    gpu_ids = opt['gpu_ids']
    device = torch.device('cuda' if gpu_ids else 'cpu')
    feature_layer = 49 if use_bn else 34
    if 'arch' in kwargs.keys() and 'vgg' in kwargs['arch']:
        if len(kwargs['arch']) > len('vgg11_'):
            feature_layer = int(kwargs['arch'][len('vgg11_'):])
        kwargs['arch'] = kwargs['arch'][:len('vgg11')]
    netF = arch.VGGFeatureExtractor(feature_layer=feature_layer, use_bn=use_bn, use_input_norm=True, device=device, **kwargs)
    if gpu_ids:
        netF = SingleDeviceDataParallel(netF.to(device))
    netF.eval()  # No need to train
    return netF
