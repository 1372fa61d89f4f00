"""Generator architecture with the reference's constructor signature, attributes and state-dict keys
(models/modules/architecture.py:228-302), executed by esr_b200.engine.RRDBEngine."""
import math

import torch
import torch.nn as nn

from . import block as B


class RRDBNet(nn.Module):
    def __init__(self, in_nc, out_nc, nf, nb, gc=32, upscale=4, norm_type=None, act_type='leakyrelu', mode='CNA',
                 upsample_mode='upconv', latent_input=None, num_latent_channels=None):
        super(RRDBNet, self).__init__()
        self.latent_input = None
        if num_latent_channels is not None and num_latent_channels > 0:
            self.latent_input = latent_input
            num_latent_channels_HR = 1 * num_latent_channels
            if 'HR_rearranged' in latent_input:
                num_latent_channels *= upscale ** 2
        self.num_latent_channels = 1 * num_latent_channels
        self.upscale = upscale
        n_upscale = 1 if upscale == 3 else int(math.log(upscale, 2))
        if latent_input is not None:
            in_nc += num_latent_channels
        if latent_input is None or 'all_layers' not in latent_input:
            num_latent_channels, num_latent_channels_HR = 0, 0
        if act_type != 'leakyrelu':
            raise NotImplementedError('esr_b200 RRDBNet: only leakyrelu(0.2) is built')
        if upsample_mode not in ('upconv', 'pixelshuffle'):
            raise NotImplementedError('upsample mode [{:s}] is not found'.format(upsample_mode))
        if self.latent_input is not None and ('HR_downscaled' not in self.latent_input):
            # the reference itself only works in this domain ('LR' fails at construction, 'HR_rearranged' raises)
            raise NotImplementedError('latent_input domain must be HR_downscaled')
        # engine-facing description.  NOTE: like the reference (architecture.py:250) the dense blocks are
        # built with a literal growth of 32 whatever `gc` says.
        self.nf, self.gc, self.out_nc, self.norm_type, self.upsample_mode = nf, 32, out_nc, norm_type, upsample_mode
        self.z_lead = num_latent_channels
        self._n_upscale = n_upscale

        fea_conv = B.conv_block(in_nc, nf, kernel_size=3, norm_type=None, act_type=None, return_module_list=True)
        rb_blocks = [B.RRDB(nf, kernel_size=3, gc=32, stride=1, bias=True, pad_type='zero', norm_type=norm_type, act_type=act_type,
                            mode='CNA', latent_input_channels=num_latent_channels) for _ in range(nb)]
        LR_conv = B.conv_block(nf + num_latent_channels, nf, kernel_size=3, norm_type=norm_type, act_type=None, mode=mode,
                               return_module_list=True)
        upsample_block = B.upconv_blcok if upsample_mode == 'upconv' else B.pixelshuffle_block
        if upscale == 3:
            upsampler = [upsample_block(nf, nf, 3, act_type=act_type)]
        else:
            upsampler = [upsample_block(nf, nf, act_type=act_type) for _ in range(n_upscale)]
        HR_conv0 = B.conv_block(nf + num_latent_channels_HR, nf, kernel_size=3, norm_type=None, act_type=act_type, return_module_list=True)
        HR_conv1 = B.conv_block(nf + num_latent_channels_HR, out_nc, kernel_size=3, norm_type=None, act_type=None, return_module_list=True)
        as_list = lambda m: m if isinstance(m, list) else [m]
        shortcut = B.ShortcutBlock(rb_blocks + as_list(LR_conv), latent_input_channels=num_latent_channels, use_module_list=True)
        self.model = nn.ModuleList(as_list(fea_conv) + [shortcut] + upsampler + as_list(HR_conv0) + as_list(HR_conv1))
        self._engines = {}
        self.compute_dtype = torch.float16

    # ---- engine-facing helpers ----------------------------------------------------------------------
    def upsamplers(self):
        return [self.model[2 + k] for k in range(self._n_upscale)]

    def up_factors(self):
        return [3] if self.upscale == 3 else [2] * self._n_upscale

    def engine(self, dtype=None, training=False):
        """engine for `dtype` (default: self.compute_dtype, fp16).  Training (weight gradients) uses the bf16 engine:
        fp16 gradients underflow, and the tensor cores want activations, weights and gradients in one format.  In the
        'parity' arithmetic mode (esr_b200.precision) every pass runs the split-precision engine."""
        from esr_b200 import ops, precision
        from esr_b200.engine import RRDBEngine
        if dtype is not None:
            key = dtype
        elif precision.parity():
            key = ops.SPLIT
        else:
            key = torch.bfloat16 if training else self.compute_dtype
        if key not in self._engines:
            self._engines[key] = RRDBEngine(self, dtype=key)
        return self._engines[key]

    def forward(self, x, pad=0):
        """x: [N, z*s^2 + in_nc, h, w] exactly as the reference feeds it (latent packed by
        SRRaGANModel.Prepare_Input).  `pad` > 0 replicate-pads the input (CEM eval mode) inside the
        packing kernel."""
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from esr_b200.autograd import rrdb_forward_with_grad
            return rrdb_forward_with_grad(self, x, pad)
        # a model under training feeds its critic from no-grad forwards too (D-only steps): same precision as its training passes
        return self.engine(training=bool(getattr(self, 'train_precision_always', False) and self.training)).forward(x, pad=pad)


class VGGFeatureExtractor(nn.Module):
    """Perceptual-loss feature extractor with the reference's constructor and state-dict keys
    (models/modules/architecture.py:658-724): torchvision's vgg19().features[:feature_layer + 1] behind the ImageNet
    input normalisation, frozen.  Runs on esr_b200.vgg.VGGEngine (fused conv launches + max-pool kernels).
    torchvision's pretrained weights cannot be downloaded here: pass `state_dict` (torchvision key layout
    `features.N.weight`, optionally prefixed `module.`), or load one later; otherwise the weights are torchvision's
    own initialisation (kaiming-normal fan_out), which is what `arch_config='untrained'` asks for in the reference."""

    def __init__(self, feature_layer=34, use_bn=False, use_input_norm=True, device=torch.device('cpu'), state_dict=None, arch='vgg19',
                 arch_config='', **kwargs):
        super(VGGFeatureExtractor, self).__init__()
        if use_bn or arch != 'vgg19' or arch_config.replace('untrained_', '').replace('untrained', '') != '':
            raise NotImplementedError('esr_b200 VGGFeatureExtractor: only plain vgg19 (no batch norm, no modified architecture) is built')
        from esr_b200.vgg import vgg19_layers
        self.feature_layer = feature_layer
        mods = []
        for kind, idx, cin, cout in vgg19_layers(feature_layer):
            mods.append(nn.Conv2d(cin, cout, 3, padding=1) if kind == 'conv' else (nn.ReLU(inplace=True) if kind == 'relu' else nn.MaxPool2d(2, 2)))
        self.features = nn.Sequential(*mods)
        for m in self.features.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                nn.init.constant_(m.bias, 0)
        if state_dict is not None:
            state_dict = dict(zip([key.replace('module.', '') for key in state_dict.keys()], [value for value in state_dict.values()]))
            self.load_state_dict({k: v for k, v in state_dict.items() if k in self.state_dict()}, strict=False)
        elif 'untrained' not in arch_config:
            print('WARNING: pretrained VGG19 weights are not available offline; VGGFeatureExtractor starts from torchvision\'s random '
                  'initialisation until a state_dict is loaded.')
        self.use_input_norm = use_input_norm
        if self.use_input_norm:
            self.register_buffer('mean', torch.Tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1).to(device))
            self.register_buffer('std', torch.Tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1).to(device))
        for k, v in self.features.named_parameters():
            v.requires_grad = False
        self.compute_dtype = torch.float16    # backward uses loss scaling (esr_b200.vgg): fp16's 10-bit mantissa, bf16's range not needed
        self._engines = {}

    def engine(self):
        from esr_b200 import ops, precision
        from esr_b200.vgg import VGGEngine
        key = ops.SPLIT if precision.parity() else self.compute_dtype
        if key not in self._engines:
            self._engines[key] = VGGEngine(self, dtype=key)
        return self._engines[key]

    def forward(self, x):
        from esr_b200.vgg import vgg_forward
        return vgg_forward(self, x)


class Discriminator_VGG_128(nn.Module):
    """VGG-style critic with the reference's constructor, module tree and state-dict keys (`features.{0,2,3,5,6,...}`,
    `classifier.{0,2}`; models/modules/architecture.py:446-508): conv0 3x3 (no norm) then alternating 4x4 stride-2 / 3x3
    convs up to 8*base_nf channels, each Conv -> BatchNorm2d -> LeakyReLU(0.2), then Linear(8*base_nf*s*s -> 100) -> LeakyReLU
    -> Linear(100 -> 1).  Runs on esr_b200.disc.DiscEngine (tensor-core conv launches, BatchNorm / Linear kernels).  The
    truncated (`nb` < 10) and patch-discriminator (`num_2_strides` < 5) variants are not built."""

    def __init__(self, in_nc, base_nf, norm_type='batch', act_type='leakyrelu', mode='CNA', input_patch_size=128, num_2_strides=5, nb=10):
        super(Discriminator_VGG_128, self).__init__()
        assert num_2_strides <= 5, 'Can be modified by adding more stridable layers, if needed.'
        nb = 10 if nb is None else nb
        if num_2_strides != 5 or nb < 10:
            raise NotImplementedError('esr_b200 Discriminator_VGG_128: only the full 10-layer, 5-stride network with the FC classifier is built')
        if norm_type != 'batch' or act_type != 'leakyrelu' or mode != 'CNA':
            raise NotImplementedError('esr_b200 Discriminator_VGG_128: only batch norm + leakyrelu(0.2) in CNA order is built')
        self.num_2_strides = 5
        if int(input_patch_size) % 32 != 0:
            # every 4x4 stride-2 conv runs over the 2x2 space-to-depth image: its input must be even, five times over (the reference
            # sizes its classifier with ceil((s - 1) / 2) and accepts odd sizes, architecture.py:457-487)
            raise NotImplementedError('esr_b200 Discriminator_VGG_128: input_patch_size must be a multiple of 32 (got %d); with the CEM the '
                                      'critic sees patch_size - 2 * invalidity_margins_HR pixels' % int(input_patch_size))
        size = 1 * input_patch_size
        chans = [(in_nc, base_nf, 3), (base_nf, base_nf, 4), (base_nf, base_nf * 2, 3), (base_nf * 2, base_nf * 2, 4),
                 (base_nf * 2, base_nf * 4, 3), (base_nf * 4, base_nf * 4, 4), (base_nf * 4, base_nf * 8, 3), (base_nf * 8, base_nf * 8, 4),
                 (base_nf * 8, base_nf * 8, 3), (base_nf * 8, base_nf * 8, 4)]
        blocks = []
        for k, (ci, co, ks) in enumerate(chans):
            blocks.append(B.conv_block(ci, co, kernel_size=ks, stride=2 if ks == 4 else 1, norm_type=None if k == 0 else norm_type,
                                       act_type=act_type, mode=mode))
            if ks == 4:
                size = math.ceil((size - 1) / 2)
        self.features = B.sequential(*blocks)
        self.last_FC_layers = True
        self.classifier = nn.Sequential(nn.Linear(base_nf * 8 * int(size) ** 2, 100), nn.LeakyReLU(0.2, True), nn.Linear(100, 1))
        self.compute_dtype = torch.bfloat16   # training precision of the step (BASELINE config 3); fp16 is a switch for inference use
        self.supports_double_backward = True   # WGAN-GP's second-order pass: esr_b200.disc.DiscEngine.second_order_param_grads
        self._engines = {}

    def engine(self):
        from esr_b200 import ops, precision
        from esr_b200.disc import DiscEngine
        key = ops.SPLIT if precision.parity() else self.compute_dtype
        if key not in self._engines:
            self._engines[key] = DiscEngine(self, dtype=key)
        return self._engines[key]

    def forward(self, x):
        from esr_b200.disc import disc_forward
        return disc_forward(self, x)
