"""Loss helpers with the reference's names (models/modules/loss.py).  Only what the generator / Z-optimisation paths
construct is built: `Latent_channels_desc_2_num_channels`, `FilterLoss` (structure-tensor descriptors: one CUDA pass per
image batch + the reference's host-side percentile history), `GANLoss`, `CreateRangeLoss`, `GradientPenaltyLoss` (raises).
The GAN / range losses are a few reductions on [B,1] logits or one image - host-level PyTorch, as in the reference."""
import re
from collections import deque

import numpy as np
import torch
import torch.nn as nn


def Latent_channels_desc_2_num_channels(latent_channels_desc):
    if isinstance(latent_channels_desc, int):
        return latent_channels_desc
    if latent_channels_desc == 'STD_1dir':
        return 2
    if latent_channels_desc == 'STD_directional' or 'structure_tensor' in latent_channels_desc:
        m = re.search(r'(\d)+', latent_channels_desc)
        return int(m.group(0)) if m is not None else 3


class _StructureTensorFn(torch.autograd.Function):
    """[N,C,H,W] -> [N,3] per-image means of (dx^2, dy^2, dx*dy) of the 2x2 finite differences (esr_structure_tensor_fwd/_bwd)"""

    @staticmethod
    def forward(ctx, img):
        from esr_b200 import ops
        img = img.float().contiguous()
        ctx.save_for_backward(img)
        return ops.structure_tensor(img)

    @staticmethod
    def backward(ctx, g):
        from esr_b200 import ops
        img, = ctx.saved_tensors
        return ops.structure_tensor_bwd(img, g.float().contiguous())


def structure_tensor_means(img):
    return _StructureTensorFn.apply(img)


class FilterLoss(nn.Module):
    """L_struct (models/modules/loss.py:27-209): ties the latent code Z to a measurable property of the output.  Built for the
    structure-tensor descriptors of the explorable-SR configuration ('structure_tensor', 'SVDinNormedOut_structure_tensor',
    options/train/train_explorable_SR.json:45) in model-training mode: the per-image structure tensor (mean dx^2, dy^2, dx*dy of
    2x2 finite differences, :51-62,140-147) of the output, normalised by the ground truth's (:159-165), must equal the spatial
    mean of Z mapped affinely onto the running 5-95 percentile range of the measured values (:167-178; the percentile history
    lives on the host exactly as in the reference, one small device->host read per call).  Returns |measured - f(Z)| [B,3].
    The statistics run as one HBM-bound CUDA pass per image batch; the STD / SVD variants and the constant-Z (GUI) mode are
    not built."""

    def __init__(self, latent_channels, constant_Z=None, reference_images=None, masks=None, task='SR', gray_scale=False):
        super(FilterLoss, self).__init__()
        self.data_keys = {'reconstructed': 'SR', 'GT': 'HR'} if task == 'SR' else {'reconstructed': 'Decomp', 'GT': 'Uncomp'}
        self.latent_channels = latent_channels
        self.num_channels = Latent_channels_desc_2_num_channels(self.latent_channels) if latent_channels is not None else 0
        if not self.num_channels:
            self.num_channels = 0
            return
        self.NOISE_STD = 1e-15
        self.model_training = isinstance(self.latent_channels, str)
        self.built = False
        if self.model_training:
            if 'structure_tensor' in self.latent_channels and self.latent_channels != 'SVD_structure_tensor' and constant_Z is None \
                    and not gray_scale:
                self.NOISE_STD = 1 / 255 if task == 'SR' else 1
                self.built = True
            self.collected_ratios = [deque(maxlen=10000) for _ in range(self.num_channels)]

    def forward(self, data):
        if not getattr(self, 'built', False):
            raise NotImplementedError('esr_b200 FilterLoss: only the structure_tensor / SVDinNormedOut_structure_tensor descriptors in '
                                      'model-training mode are built (got %r)' % (self.latent_channels,))
        LOWER_PERCENTILE, HIGHER_PERCENTILE = 5, 95
        cur_Z = data['Z'].mean(dim=(2, 3))
        d_sr = structure_tensor_means(data[self.data_keys['reconstructed']])             # [B,3]
        with torch.no_grad():
            d_hr = structure_tensor_means(data[self.data_keys['GT']])
        if self.latent_channels == 'SVDinNormedOut_structure_tensor':                    # RATIO_LOSS 'SingleNormalizer' (:162-164)
            normalizer = torch.sqrt(d_hr[:, 0]) * torch.sqrt(d_hr[:, 1])
            measured = [d_sr[:, i] / (normalizer + self.NOISE_STD) for i in range(3)]
        else:                                                                            # 'OnlyDiagonals' (:165-167)
            measured = [d_sr[:, i] / (d_hr[:, i] + torch.sign(d_sr[:, i]) * self.NOISE_STD) if i < 2 else d_sr[:, i] / 1 for i in range(3)]
        host = torch.stack(measured, 1).detach().cpu().numpy()                           # the running percentile history is host state
        normalized_Z = []
        for i in range(len(measured)):
            self.collected_ratios[i] += [float(v) for v in host[:, i]]
            upper_bound = np.percentile(self.collected_ratios[i], HIGHER_PERCENTILE)
            lower_bound = np.percentile(self.collected_ratios[i], LOWER_PERCENTILE)
            normalized_Z.append((cur_Z[:, i]) / 2 * (upper_bound - lower_bound) + np.mean([upper_bound, lower_bound]))
        return (torch.stack(measured, 1) - torch.stack(normalized_Z, 1)).abs()


class GANLoss(nn.Module):
    """vanilla = BCE-with-logits against constant labels, lsgan = MSE, wgan* = -/+ mean (loss.py:212-246)."""

    def __init__(self, gan_type, real_label_val=1.0, fake_label_val=0.0):
        super(GANLoss, self).__init__()
        self.gan_type = gan_type.lower()
        self.real_label_val, self.fake_label_val = real_label_val, fake_label_val
        if self.gan_type == 'vanilla':
            self.loss = nn.BCEWithLogitsLoss()
        elif self.gan_type == 'lsgan':
            self.loss = nn.MSELoss()
        elif 'wgan' in self.gan_type:
            self.loss = lambda input, target: -1 * input.mean() if target else input.mean()
        else:
            raise NotImplementedError('GAN type [{:s}] is not found'.format(self.gan_type))

    def get_target_label(self, input, target_is_real):
        if 'wgan' in self.gan_type:
            return target_is_real
        return torch.empty_like(input).fill_(self.real_label_val if target_is_real else self.fake_label_val)

    def forward(self, input, target_is_real, hinge_threshold=None):
        if hinge_threshold is not None:
            input = torch.clamp_max(input, hinge_threshold) if target_is_real else torch.clamp_min(input, -1 * hinge_threshold)
        return self.loss(input, self.get_target_label(input, target_is_real))


def CreateRangeLoss(legit_range, chroma_mode=False):
    lo, hi = float(legit_range[0]), float(legit_range[1])

    def RangeLoss(x):
        if chroma_mode:
            x = x[:, 1:, ...]
        return torch.max(torch.clamp_min(x - hi, 0), torch.clamp_min(lo - x, 0)).mean()
    return RangeLoss


class GradientPenaltyLoss(nn.Module):
    """WGAN-GP penalty ((||d crit / d interp||_2 - 1)^2).mean() (loss.py:260-279).  It differentiates THROUGH the critic's input
    gradient.  For logits of esr_b200's Discriminator_VGG_128 the penalty and its parameter gradient are computed by the engine
    (esr_b200.disc._GradPenaltyFn: a tangent forward along dL/dg, then a backward over the primal / tangent pair - BatchNorm's
    double backward is its own kernel); any other critic (plain torch modules) takes autograd's create_graph route as in the reference."""

    def __init__(self, device=torch.device('cpu')):
        super(GradientPenaltyLoss, self).__init__()

    def forward(self, interp, interp_crit):
        from esr_b200.disc import gradient_penalty      # logits of the CUDA critic: tangent-forward + double-backward launches
        fused = gradient_penalty(interp_crit)
        if fused is not None:
            return fused
        grad_interp = torch.autograd.grad(outputs=interp_crit, inputs=interp, grad_outputs=torch.ones_like(interp_crit),
                                          create_graph=True, retain_graph=True, only_inputs=True)[0]
        grad_interp_norm = grad_interp.view(grad_interp.size(0), -1).norm(2, dim=1)
        return ((grad_interp_norm - 1) ** 2).mean()
