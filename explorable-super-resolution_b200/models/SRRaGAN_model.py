"""SRRaGANModel — orchestration layer with the reference's surface (models/SRRaGAN_model.py).

Built in this round: construction for inference / latent exploration (CEM + generator + checkpoint loading),
`feed_data`, `Prepare_Input`, `GetLatent`, `test`, `Output_Batch`, `get_current_visuals`, `save`/`load` — i.e. everything
`test.py`, the GUI's `Feed_n_Run_model` and `Z_optimizer.optimize` call.  `optimize_parameters` (the GAN training step:
discriminator, VGG features, weight gradients) raises NotImplementedError until those kernels exist."""
import os
import re
from collections import OrderedDict

import numpy as np
import torch

import CEM.CEMnet as CEMnet
import models.networks as networks
from models.modules.loss import FilterLoss
from .base_model import BaseModel


class SRRaGANModel(BaseModel):
    def __init__(self, opt, accumulation_steps_per_batch=1, init_Fnet=None, init_Dnet=None, **kwargs):
        super(SRRaGANModel, self).__init__(opt)
        if self.is_train:
            raise NotImplementedError('esr_b200: the SRRaGAN training step (D, VGG features, wgrad) is not built yet; '
                                      'use is_train=False (test / latent exploration)')
        self.log_path = opt['path']['log'] if opt['path'] is not None else None
        self.latent_input_domain = opt['network_G']['latent_input_domain']
        self.latent_input = opt['network_G']['latent_input'] if opt['network_G']['latent_input'] != 'None' else None
        if self.latent_input is not None:
            self.Z_size_factor = opt['scale'] if 'HR' in opt['network_G']['latent_input_domain'] else 1
            assert isinstance(opt['network_G']['latent_channels'], int)
        self.cri_latent = None
        self.optimalZ_loss_type = None
        self.num_latent_channels = FilterLoss(latent_channels=opt['network_G']['latent_channels']).num_channels
        self.CEM_net = None
        self.CEM_arch = opt['network_G']['CEM_arch']
        self.step = 0
        self.D_exists = False
        self.optimizer_G = None
        if self.CEM_arch or self.latent_input is not None:
            CEM_conf = CEMnet.Get_CEM_Conf(opt['scale'])
            CEM_conf.sigmoid_range_limit = bool(opt['network_G']['sigmoid_range_limit'])
            CEM_conf.input_range = np.array(opt['range']) if opt['range'] is not None else np.array([0, 1])
            if opt['test'] is not None and opt['test']['kernel'] == 'estimated':
                CEM_conf.lower_magnitude_bound = 0.1
            kernel = kwargs['kernel'] if 'kernel' in kwargs.keys() else (None if opt['test'] is None else opt['test']['kernel'])
            self.CEM_net = CEMnet.CEMnet(CEM_conf, upscale_kernel=kernel)
            if not self.CEM_arch:
                self.CEM_net.WrapArchitecture_PyTorch(only_padders=True)
        opt['network_G']['scale'] = opt['network_G']['scale'] if opt['network_G']['scale'] is not None else opt['scale']
        self.netG = networks.define_G(opt, CEM=self.CEM_net, num_latent_channels=self.num_latent_channels)
        self.netG.to(self.device)
        self.netG.eval()
        self.Set_Require_Grad_Status(self.netG, False)
        self.load()

    # ---- I/O of one batch -------------------------------------------------------------------------
    def Output_Batch(self, within_0_1):
        return torch.clamp(self.fake_H, 0, 1) if within_0_1 else self.fake_H

    def Prepare_Input(self, LR_image, latent_input, **kwargs):
        """Z[B,c,sH,sW] is re-viewed (raw memory, not a pixel-unshuffle) as [B, c*s^2, H, W] and concatenated in front of LR."""
        if latent_input is not None:
            if LR_image.size()[2:] != latent_input.size()[2:]:
                latent_input = latent_input.contiguous().view([latent_input.size(0)] + [latent_input.size(1) * self.opt['scale'] ** 2] +
                                                              list(LR_image.size()[2:]))
            self.model_input = torch.cat([latent_input, LR_image], dim=1)
        else:
            self.model_input = 1 * LR_image

    def GetLatent(self):
        latent = 1 * self.model_input[:, :-3, ...]
        if latent.size(1) != self.num_latent_channels:
            latent = latent.view([latent.size(0)] + [self.num_latent_channels] + [self.opt['scale'] * v for v in list(latent.size()[2:])])
        return latent

    def feed_data(self, data, need_GT=True, **kwargs):
        self.var_L = data['LR'].to(self.device)
        cur_Z = None
        if self.latent_input is not None:
            hr_size = [self.Z_size_factor * v for v in list(self.var_L.size()[2:])]
            if 'Z' in data.keys():
                cur_Z = data['Z']
            else:
                cur_Z = 2 * torch.rand([self.var_L.size(0), self.num_latent_channels] + hr_size).type(self.var_L.type()) - 1
            if isinstance(cur_Z, (int, float)) or (not torch.is_tensor(cur_Z) and (np.ndim(cur_Z) < 4 or np.shape(cur_Z)[2] == 1)):
                cur_Z = cur_Z * np.ones([1, self.num_latent_channels] + hr_size)
            elif torch.is_tensor(cur_Z) and cur_Z.dim() == 4 and cur_Z.size(2) == 1:
                cur_Z = (cur_Z * torch.ones([1, 1] + hr_size, device=cur_Z.device)).type(self.var_L.type())
            if not torch.is_tensor(cur_Z):
                cur_Z = torch.from_numpy(np.asarray(cur_Z)).type(self.var_L.type())
            cur_Z = cur_Z.to(self.device)
            if cur_Z.size(0) == 1 and self.var_L.size(0) > 1:
                cur_Z = cur_Z.expand([self.var_L.size(0)] + list(cur_Z.shape[1:]))
        self.Prepare_Input(LR_image=self.var_L, latent_input=cur_Z)
        if need_GT:
            self.var_H = data['HR'].to(self.device)
            self.var_ref = (data['ref'] if 'ref' in data else data['HR']).to(self.device)

    def optimize_parameters(self):
        raise NotImplementedError('esr_b200: SRRaGANModel.optimize_parameters is not built yet (SURVEY 8a-12..16)')

    def test(self, prevent_grads_calc=True, **kwargs):
        self.netG.eval()
        if prevent_grads_calc:
            with torch.no_grad():
                self.fake_H = self.netG(self.model_input)
        else:
            self.fake_H = self.netG(self.model_input)
        self.output_image = 1 * self.fake_H
        self.netG.train()

    def get_current_visuals(self, need_HR=True, entire_batch=False, to_cpu=True):
        sel = (lambda t: t.detach().float()) if entire_batch else (lambda t: t.detach()[0].float())
        out = OrderedDict()
        out['LR'] = sel(self.var_L)
        out['SR'] = sel(self.fake_H)
        if need_HR:
            out['HR'] = sel(self.var_H)
        if to_cpu:
            for k in out:
                out[k] = out[k].cpu()
        return out

    def get_current_log(self):
        return {}

    # ---- checkpoints --------------------------------------------------------------------------------
    def load(self, max_step=None, resume_train=None):
        path = self.opt['path']
        models_dir = path['models'] if path is not None else None
        own = [n for n in os.listdir(models_dir) if '_G.pth' in n] if (models_dir and os.path.isdir(models_dir)) else []
        if own:
            step_of = lambda n: int(re.search(r'(\d)+(?=_G.pth)', n).group(0))
            own = sorted(own, key=step_of)
            if max_step is not None:
                own = [n for n in own if step_of(n) <= max_step]
            name = own[-1]
            print('Testing model for G [{:s}] ...'.format(os.path.join(models_dir, name)))
            self.load_network(os.path.join(models_dir, name), self.netG)
            self.gradient_step_num = step_of(name)
        elif path is not None and path['pretrained_model_G'] is not None:
            print('loading model for G [{:s}] ...'.format(path['pretrained_model_G']))
            self.load_network(path['pretrained_model_G'], self.netG)

    def save(self, iter_label):
        return self.save_network(self.save_dir, self.netG, 'G', iter_label, self.optimizer_G)
