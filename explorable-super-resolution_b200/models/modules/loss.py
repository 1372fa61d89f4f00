"""Loss helpers with the reference's names (models/modules/loss.py).  Only what the generator / Z-optimisation paths
construct is built: `Latent_channels_desc_2_num_channels`, `FilterLoss` (channel bookkeeping), `GANLoss`,
`CreateRangeLoss`, `GradientPenaltyLoss`.  They are a few reductions on [B,1] logits or one image — host-level PyTorch,
as in the reference; the structure-tensor statistics of FilterLoss.forward belong to the training step (not built)."""
import re

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


class FilterLoss(nn.Module):
    def __init__(self, latent_channels, constant_Z=None, reference_images=None, masks=None, task='SR', gray_scale=False):
        super(FilterLoss, self).__init__()
        self.latent_channels = latent_channels
        self.num_channels = Latent_channels_desc_2_num_channels(self.latent_channels) if latent_channels is not None else 0
        if not self.num_channels:
            self.num_channels = 0

    def forward(self, data):
        raise NotImplementedError('esr_b200: FilterLoss.forward (L_struct of the training step, SURVEY 8a-15) is not built yet')


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
    def __init__(self, device=torch.device('cpu')):
        super(GradientPenaltyLoss, self).__init__()

    def forward(self, interp, interp_crit):
        raise NotImplementedError('esr_b200: WGAN-GP needs a double backward through the discriminator (SURVEY 8f-2), not built yet')
