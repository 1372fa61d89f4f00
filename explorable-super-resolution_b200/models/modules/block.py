"""Building blocks of the generator — parameter containers with the reference's names.

The module tree (attribute names, nesting, ModuleList/Sequential indices) reproduces
models/modules/block.py of the reference so that `state_dict()` keys are identical
(`...RDB1.convs.0.0.weight`, `model.1.sub.23.weight`, ...).  Unlike the reference these classes do not
compute anything themselves: the whole generator runs as fused CUDA launches scheduled by
esr_b200.engine.RRDBEngine, so calling a block's forward() directly raises (there is no eager path)."""
import torch.nn as nn


def _no_eager(name):
    raise NotImplementedError('%s has no stand-alone forward in esr_b200: the generator runs through '
                              'RRDBNet.forward (fused CUDA launches); there is no eager/CPU path' % name)


def act(act_type, inplace=True, neg_slope=0.2, n_prelu=1):
    kind = act_type.lower()
    if kind == 'relu':
        return nn.ReLU(inplace)
    if kind == 'leakyrelu':
        return nn.LeakyReLU(neg_slope, inplace)
    if kind == 'prelu':
        return nn.PReLU(num_parameters=n_prelu, init=neg_slope)
    raise NotImplementedError('activation layer [{:s}] is not found'.format(kind))


def norm(norm_type, nc):
    kind = norm_type.lower()
    if kind == 'batch':
        return nn.BatchNorm2d(nc, affine=True)
    if kind == 'instance':
        return nn.InstanceNorm2d(nc, affine=False)
    raise NotImplementedError('normalization layer [{:s}] is not found'.format(kind))


def pad(pad_type, padding):
    kind = pad_type.lower()
    if padding == 0:
        return None
    if kind == 'reflect':
        return nn.ReflectionPad2d(padding)
    if kind == 'replicate':
        return nn.ReplicationPad2d(padding)
    raise NotImplementedError('padding layer [{:s}] is not implemented'.format(kind))


def get_valid_padding(kernel_size, dilation):
    return (kernel_size + (kernel_size - 1) * (dilation - 1) - 1) // 2


def sequential(*args, return_module_list=False):
    """Flattens nested nn.Sequential and drops None entries; a single argument is returned as is."""
    if len(args) == 1:
        return args[0]
    flat = []
    for m in args:
        if isinstance(m, nn.Sequential):
            flat.extend(m.children())
        elif isinstance(m, nn.Module):
            flat.append(m)
    return flat if return_module_list else nn.Sequential(*flat)


def conv_block(in_nc, out_nc, kernel_size, stride=1, dilation=1, groups=1, bias=True, pad_type='zero', norm_type=None,
               act_type='relu', mode='CNA', return_module_list=False):
    """Conv (+norm) (+act) in CNA order (NAC: norm, act, conv)."""
    assert mode in ['CNA', 'NAC', 'CNAC'], 'Wong conv mode [{:s}]'.format(mode)
    padding = get_valid_padding(kernel_size, dilation)
    p = pad(pad_type, padding) if pad_type and pad_type != 'zero' else None
    c = nn.Conv2d(in_nc, out_nc, kernel_size=kernel_size, stride=stride, padding=padding if pad_type == 'zero' else 0,
                  dilation=dilation, bias=bias, groups=groups)
    a = act(act_type) if act_type else None
    if 'CNA' in mode:
        n = norm(norm_type, out_nc) if norm_type else None
        return sequential(p, c, n, a, return_module_list=return_module_list)
    if norm_type is None and act_type is not None:
        a = act(act_type, inplace=False)
    n = norm(norm_type, in_nc) if norm_type else None
    return sequential(n, a, p, c)


class ShortcutBlock(nn.Module):
    """x[:, z:] + sub(x) (block.py:76-97 of the reference); container only."""

    def __init__(self, submodule, latent_input_channels=0, use_module_list=False):
        super(ShortcutBlock, self).__init__()
        self.sub = nn.ModuleList(submodule) if use_module_list else submodule
        self.num_latent_channels = latent_input_channels

    def forward(self, x):
        _no_eager('ShortcutBlock')


class ResidualDenseBlock_5C(nn.Module):
    """Five 3x3 convs on a growing channel stack; returns x5*0.2 + x (block.py:196-235); container only."""

    def __init__(self, nc, kernel_size=3, gc=32, stride=1, bias=True, pad_type='zero', norm_type=None, act_type='leakyrelu',
                 mode='CNA', latent_input_channels=0):
        super(ResidualDenseBlock_5C, self).__init__()
        self.USE_MODULE_LIST = True
        last_act = None if mode == 'CNA' else act_type
        self.convs = nn.ModuleList([
            conv_block(nc + i * gc + latent_input_channels, gc if i < 4 else nc, kernel_size if i < 4 else 3, stride, bias=bias,
                       pad_type=pad_type, norm_type=norm_type, act_type=act_type if i < 4 else last_act, mode=mode)
            for i in range(5)])
        # a lone conv comes back bare from conv_block; the reference keeps it inside a Sequential (key `convs.4.0`)
        for i, m in enumerate(self.convs):
            if isinstance(m, nn.Conv2d):
                self.convs[i] = nn.Sequential(m)

    def forward(self, x):
        _no_eager('ResidualDenseBlock_5C')


class RRDB(nn.Module):
    """RDB3(RDB2(RDB1(x)))*0.2 + x (block.py:245-270); container only."""

    def __init__(self, nc, kernel_size=3, gc=32, stride=1, bias=True, pad_type='zero', norm_type=None, act_type='leakyrelu',
                 mode='CNA', latent_input_channels=0):
        super(RRDB, self).__init__()
        self.num_latent_channels = latent_input_channels
        mk = lambda: ResidualDenseBlock_5C(nc, kernel_size, gc, stride, bias, pad_type, norm_type, act_type, mode,
                                           latent_input_channels)
        self.RDB1, self.RDB2, self.RDB3 = mk(), mk(), mk()

    def forward(self, x):
        _no_eager('RRDB')


class Upsampler(nn.Module):
    def __init__(self, upscale_factor, mode):
        super(Upsampler, self).__init__()
        self.upscale_factor = upscale_factor
        self.mode = mode

    def forward(self, input):
        _no_eager('Upsampler')


def pixelshuffle_block(in_nc, out_nc, upscale_factor=2, kernel_size=3, stride=1, bias=True, pad_type='zero', norm_type=None,
                       act_type='relu'):
    """conv(in -> out*r^2) -> PixelShuffle(r) -> (norm) -> (act)."""
    conv = conv_block(in_nc, out_nc * (upscale_factor ** 2), kernel_size, stride, bias=bias, pad_type=pad_type, norm_type=None,
                      act_type=None)
    n = norm(norm_type, out_nc) if norm_type else None
    a = act(act_type) if act_type else None
    return sequential(conv, nn.PixelShuffle(upscale_factor), n, a)


def upconv_blcok(in_nc, out_nc, upscale_factor=2, kernel_size=3, stride=1, bias=True, pad_type='zero', norm_type=None,
                 act_type='relu', mode='nearest'):
    """nearest up-sampling followed by conv (+act)."""
    conv = conv_block(in_nc, out_nc, kernel_size, stride, bias=bias, pad_type=pad_type, norm_type=norm_type, act_type=act_type)
    return sequential(Upsampler(upscale_factor, mode), conv)
