"""SRRaGANModel — orchestration layer with the reference's surface (models/SRRaGAN_model.py).

Built: construction for inference / latent exploration (CEM + generator + checkpoint loading), `feed_data`,
`Prepare_Input`, `GetLatent`, `test`, `Output_Batch`, `get_current_visuals`, `save`/`load` — everything `test.py`, the GUI's
`Feed_n_Run_model` and `Z_optimizer.optimize` call — and the training step (`optimize_parameters`,
models/SRRaGAN_model.py:280-519): discriminator step (Discriminator_VGG_128, vanilla / lsgan / wgan losses, relativistic by
default) and generator step (pixel + VGG-feature + range + GAN + latent-control losses), gradient accumulation for both, Adam, MultiStepLR,
D_update_ratio / D_init_iters scheduling, D verification ('past' / 'current' / 'convergence') and the optimised-Z reference loss
L_map.  Configurations that need the decomposed-output critic, the automatic D_update_ratio controller or the histogram
reference loss raise NotImplementedError at construction: those are not built and there is no PyTorch fallback."""
import os
import re
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
from torch.optim import lr_scheduler

import CEM.CEMnet as CEMnet
import models.networks as networks
from esr_b200 import parallel
from esr_b200.losses import L1Loss as EsrL1Loss, relativistic_bce
from esr_b200.optim import FlatAdam
from models.modules.loss import FilterLoss, CreateRangeLoss, GANLoss, GradientPenaltyLoss
from .base_model import BaseModel


def SVD_2_LatentZ(SVD_values, max_lambda=1):
    """utils/util.py:285-291: (lambda0, lambda1, theta) -> (sum Ix^2, sum Iy^2, sum Ix*Iy) mapped to a symmetric range"""
    l0, l1, th = SVD_values[:, 0, ...], SVD_values[:, 1, ...], SVD_values[:, -1, ...]
    return torch.stack([2 * max_lambda * (l1 * (torch.sin(th) ** 2) + l0 * (torch.cos(th) ** 2)) - max_lambda,
                        2 * max_lambda * (l0 * (torch.sin(th) ** 2) + l1 * (torch.cos(th) ** 2)) - max_lambda,
                        2 * (l0 - l1) * torch.sin(th) * torch.cos(th)], 1)


def _tensor2img(t, min_max=(0, 1)):
    """utils/util.py:196-228 for one image, float output: [C,H,W] RGB tensor -> [H,W,C] BGR array clipped and scaled to [0,1]"""
    img = t.squeeze().float().cpu().numpy()
    if img.ndim == 3:
        img = np.transpose(img[[2, 1, 0], :, :], (1, 2, 0))
    img = (np.clip(img, min_max[0], min_max[1]) - min_max[0]) / (min_max[1] - min_max[0])
    return img.astype(np.float32)


def _psnr(img1, img2):
    """utils/util.py:340-347, images in [0,255]"""
    mse = np.mean((img1.astype(np.float64) - img2.astype(np.float64)) ** 2)
    return float('inf') if mse == 0 else 20 * np.log10(255.0 / np.sqrt(mse))


def _save_png(img, path):
    """utils/util.py:231-232 (cv2.imwrite of a BGR uint8 array).  Without OpenCV the same image is written by a minimal PNG encoder
    (8-bit RGB / grey, no filtering): validation collages do not depend on an optional package."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    try:
        import cv2
        cv2.imwrite(path, img)
        return
    except ImportError:
        pass
    import struct
    import zlib
    a = np.ascontiguousarray(np.clip(np.asarray(img), 0, 255).astype(np.uint8))
    if a.ndim == 3 and a.shape[2] == 1:
        a = a[:, :, 0]
    if a.ndim == 3:
        a = np.ascontiguousarray(a[:, :, 2::-1])          # BGR -> RGB
    h, w = a.shape[:2]
    colour_type = 2 if a.ndim == 3 else 0
    raw = b''.join(b'\x00' + a[y].tobytes() for y in range(h))

    def chunk(tag, data):
        return struct.pack('>I', len(data)) + tag + data + struct.pack('>I', zlib.crc32(tag + data) & 0xffffffff)
    with open(path, 'wb') as f:
        f.write(b'\x89PNG\r\n\x1a\n' + chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, 8, colour_type, 0, 0, 0)) +
                chunk(b'IDAT', zlib.compress(raw, 6)) + chunk(b'IEND', b''))


class SRRaGANModel(BaseModel):
    def __init__(self, opt, accumulation_steps_per_batch=1, init_Fnet=None, init_Dnet=None, **kwargs):
        super(SRRaGANModel, self).__init__(opt)
        train_opt = opt['train'] if self.is_train else None
        if self.is_train:
            unbuilt = ['optimalZ_loss_type hist (SoftHistogramLoss)'] if (train_opt['optimalZ_loss_weight'] is not None and train_opt['optimalZ_loss_type'] == 'hist') else []
            if train_opt['gan_weight'] is not None:
                if isinstance(train_opt['D_update_ratio'], list):
                    unbuilt.append('automatic D_update_ratio controller')
                if (opt['network_D'] or {}).get('decomposed_input'):
                    unbuilt.append('decomposed_input')
            if unbuilt:
                raise NotImplementedError('esr_b200: the training step is built for the pixel / feature / range / GAN losses; %s are '
                                          'not built' % ', '.join(unbuilt))
        self.log_path = opt['path']['log'] if opt['path'] is not None else None
        self.latent_input_domain = opt['network_G']['latent_input_domain']
        self.latent_input = opt['network_G']['latent_input'] if opt['network_G']['latent_input'] != 'None' else None
        if self.latent_input is not None:
            self.Z_size_factor = opt['scale'] if 'HR' in opt['network_G']['latent_input_domain'] else 1
        self.cri_latent = None
        self.optimalZ_loss_type = None
        self.num_latent_channels = FilterLoss(latent_channels=opt['network_G']['latent_channels']).num_channels
        self.cri_optimalZ = None
        if self.latent_input is not None and self.is_train:   # latent-control loss L_struct (SRRaGAN_model.py:37-40)
            if train_opt['optimalZ_loss_type'] is not None and train_opt['optimalZ_loss_weight'] is not None:
                self.optimalZ_loss_type = train_opt['optimalZ_loss_type']      # (:68-70)
            self.l_latent_w = train_opt['latent_weight']
            if self.l_latent_w is not None:
                self.cri_latent = FilterLoss(latent_channels=opt['network_G']['latent_channels'])
                if not self.cri_latent.built:
                    raise NotImplementedError('esr_b200: latent_weight needs a structure-tensor latent descriptor (got %r)'
                                              % (opt['network_G']['latent_channels'],))
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
        logs_2_keep = ['l_g_pix', 'l_g_fea', 'l_g_range', 'l_g_gan', 'l_d_real', 'l_d_fake', 'l_d_real_fake', 'D_real', 'D_fake', 'D_logits_diff',
                       'Correctly_distinguished', 'psnr_val', 'LR_decrease', 'l_g_optimalZ', 'D_loss_STD', 'l_d_gp'] + ['l_g_latent_%d' % i for i in range(self.num_latent_channels)]
        self.log_dict = OrderedDict(zip(logs_2_keep, [[] for _ in logs_2_keep]))
        if not self.is_train:
            self.netG.eval()
            self.Set_Require_Grad_Status(self.netG, False)
            if init_Fnet:      # the GUI's feature-space / adversarial Z objectives ask for these next to the generator (:199-206)
                self.netF = networks.define_F(opt, use_bn=False).to(self.device)
                self.netF.eval()
            if init_Dnet:
                self.netD = networks.define_D(opt, CEM=self.CEM_net).to(self.device)
                self.netD.eval()
            self.load()
            return
        # ---- training state (models/SRRaGAN_model.py:68-203, generator branch)
        self.max_accumulation_steps = accumulation_steps_per_batch
        for mod in self.netG.modules():      # D-only steps feed the critic from no-grad forwards: same arithmetic as the training passes
            if hasattr(mod, 'z_lead'):
                mod.train_precision_always = True
        self.grad_accumulation_steps_G = train_opt['grad_accumulation_steps_G'] or 1
        self.grad_accumulation_steps_D = train_opt['grad_accumulation_steps_D'] or 1
        self.netG.train()
        self.l_gan_w = train_opt['gan_weight']
        self.D_exists = self.l_gan_w is not None
        # D verification (SRRaGAN_model.py:71-76): the generator only steps on the adversarial term once the critic is judged good
        # enough - by its recent log ('past'), by the current batch ('current') or by the flattening of its loss ('convergence')
        self.D_verification = train_opt['D_verification']
        assert self.D_verification in ['current', 'convergence', 'past', None]
        self.D_verified, self.verified_D_saved = self.D_verification is None, self.D_verification is None
        if self.D_verification == 'convergence':
            self.D_converged = False
        if self.D_exists:
            self.relativistic_D = opt['network_D']['relativistic'] is None or bool(opt['network_D']['relativistic'])
            self.add_quantization_noise = bool(opt['network_D']['add_quantization_noise'])
            self.netD = networks.define_D(opt, CEM=self.CEM_net).to(self.device)
            self.netD.train()
            critic = self.netD.module if isinstance(self.netD, nn.DataParallel) else self.netD
            if train_opt['gan_type'] == 'wgan-gp' and not getattr(critic, 'supports_double_backward', True):
                raise NotImplementedError('esr_b200: gan_type wgan-gp needs a double backward through the critic, which the CUDA engine of '
                                          'Discriminator_VGG_128 does not provide yet (SURVEY 8f-2, DESIGN 5b-6)')
        if train_opt['pixel_weight'] is not None:
            l_pix_type = train_opt['pixel_criterion']
            if l_pix_type == 'l1':
                self.cri_pix = EsrL1Loss().to(self.device)      # esr_l1_reduce kernel on CUDA tensors
            elif l_pix_type == 'l2':
                self.cri_pix = nn.MSELoss().to(self.device)
            else:
                raise NotImplementedError('Loss type [{:s}] not recognized.'.format(l_pix_type))
            self.l_pix_w = train_opt['pixel_weight']
        else:
            print('Remove pixel loss.')
            self.cri_pix = None
        if self.optimalZ_loss_type is not None:   # reference loss after optimising the latent input, L_map (:108-123)
            from Z_optimization import Z_optimizer
            self.l_g_optimalZ_w = train_opt['optimalZ_loss_weight']
            self.Z_optimizer = Z_optimizer(objective=self.optimalZ_loss_type,
                                           Z_size=2 * [int(opt['datasets']['train']['patch_size'] / (opt['scale'] / self.Z_size_factor))], model=self,
                                           Z_range=1, max_iters=10, initial_LR=1, batch_size=opt['datasets']['train']['batch_size'],
                                           HR_unpadder=self.CEM_net.HR_unpadder)
            if self.optimalZ_loss_type == 'l2':
                self.cri_optimalZ = nn.MSELoss().to(self.device)
            elif self.optimalZ_loss_type == 'l1':
                self.cri_optimalZ = EsrL1Loss().to(self.device)
            else:
                raise NotImplementedError('Loss type [{:s}] not recognized.'.format(self.optimalZ_loss_type))
        else:
            print('Remove reference loss with optimal Z.')
        if train_opt['range_weight'] is not None:
            self.cri_range = CreateRangeLoss(opt['range'])
            self.l_range_w = train_opt['range_weight']
        else:
            print('Remove range loss.')
            self.cri_range = None
        if train_opt['feature_weight'] is not None:   # G feature loss (SRRaGAN_model.py:124-138)
            l_fea_type = train_opt['feature_criterion']
            if l_fea_type == 'l1':
                self.cri_fea = EsrL1Loss().to(self.device)
            elif l_fea_type == 'l2':
                self.cri_fea = nn.MSELoss().to(self.device)
            else:
                raise NotImplementedError('Loss type [{:s}] not recognized.'.format(l_fea_type))
            self.l_fea_w = train_opt['feature_weight']
            self.netF = networks.define_F(opt, use_bn=False, **({'state_dict': kwargs['netF_state_dict']} if 'netF_state_dict' in kwargs else {})).to(self.device)
        else:
            print('Remove feature loss.')
            self.cri_fea = None
        if self.D_exists:   # GD gan loss (SRRaGAN_model.py:150-159)
            self.cri_gan = GANLoss(train_opt['gan_type'], 1.0, 0.0).to(self.device)
            if train_opt['gan_type'] == 'wgan-gp':      # gradient penalty (:161-165)
                self.cri_gp = GradientPenaltyLoss(device=self.device).to(self.device)
                self.l_gp_w = train_opt['gp_weight']
            self.global_D_update_ratio = train_opt['D_update_ratio'] if train_opt['D_update_ratio'] is not None else 1
            self.D_init_iters = train_opt['D_init_iters'] if train_opt['D_init_iters'] else 0
        else:
            print('Remove GAN loss')
            self.cri_gan, self.D_init_iters, self.global_D_update_ratio = None, 0, 1
        wd_G = train_opt['weight_decay_G'] if train_opt['weight_decay_G'] else 0
        optim_params = []
        for k, v in self.netG.named_parameters():
            if v.requires_grad:
                optim_params.append(v)
            else:
                print('WARNING: params [{:s}] will not optimize.'.format(k))
        self.lr_G = train_opt['lr_G']
        # torch.optim.Adam's arithmetic and state-dict layout, one fused launch per step (esr_b200.optim)
        self.optimizer_G = FlatAdam(optim_params, lr=self.lr_G, weight_decay=wd_G,
                                    betas=(train_opt['beta1_G'], train_opt['beta2_G'] if train_opt['beta2_G'] is not None else 0.999))
        self.optimizers.append(self.optimizer_G)
        self.optimizer_D = None
        if self.D_exists:
            wd_D = train_opt['weight_decay_D'] if train_opt['weight_decay_D'] else 0
            self.lr_D = train_opt['lr_D']
            self.optimizer_D = FlatAdam(self.netD.parameters(), lr=self.lr_D, weight_decay=wd_D,
                                        betas=(train_opt['beta1_D'], train_opt['beta2_D'] if train_opt['beta2_D'] is not None else 0.999))
            self.optimizers.append(self.optimizer_D)
        if train_opt['lr_scheme'] == 'MultiStepLR':
            for optimizer in self.optimizers:
                self.schedulers.append(lr_scheduler.MultiStepLR(optimizer, train_opt['lr_steps'], train_opt['lr_gamma']))
        else:
            raise NotImplementedError('MultiStepLR learning rate scheme is enough.')
        self.generator_step, self.generator_changed, self.generator_started_learning = False, True, False
        self.load()
        # learning rates after loading (SRRaGAN_model.py:208-218): once the adversarial term is in use (D verified, the default
        # without D_verification) the generator runs at the DISCRIMINATOR's learning rate
        if self.D_exists:
            self.lr_D = float(self.lr_D)
            for param_group in self.optimizer_D.param_groups:
                param_group['lr'] = self.lr_D
            if self.verified_D_saved:
                self.lr_G = 1 * self.lr_D
                if 'Z_optimizer' in self.__dict__:      # ... and a different number of Z iterations for the L_map step (:215-216)
                    self.Z_optimizer.max_iters = self.opt['train']['Num_Z_iterations'][-1]
        self.lr_G = float(self.lr_G)
        for param_group in self.optimizer_G.param_groups:
            param_group['lr'] = self.lr_G
        # one process per GPU: every replica starts from rank 0's weights (nn.DataParallel's per-forward replication, networks.py:122)
        parallel.broadcast_parameters(self.netG)
        if self.D_exists:
            parallel.broadcast_parameters(self.netD)
        if self.device.type == 'cuda':      # parameters, gradients and Adam moments move into flat buffers; the engines write gradients there
            for optimizer in self.optimizers:
                optimizer.register()

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

    # ---------------------------------------------------------------- host <-> device traffic of a training step
    def _to_device(self, t):
        """`t.to(self.device)`.  A PINNED host batch is copied on a dedicated stream: the next step's images then travel while the
        previous step still computes (the compute stream only waits for the copy's event).  ESR_SYNC_INPUT=1 keeps the blocking copy."""
        if (not torch.is_tensor(t)) or t.is_cuda or self.device.type != 'cuda' or not t.is_pinned() or os.environ.get('ESR_SYNC_INPUT', '0') == '1':
            return t.to(self.device)
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy_stream):
            d = t.to(self.device, non_blocking=True)
        cur.wait_stream(self._copy_stream)
        d.record_stream(cur)
        return d

    def _resolve_logs(self, block):
        """the logged scalars of the generator step come back through a pinned buffer and an event instead of a blocking .item():
        append the ones that have arrived (all of them when `block`) to their lists, in order"""
        pend = self.__dict__.get('_pending_logs')
        while pend:
            rec = pend[0]
            if not rec['ev'].query():
                if not block and len(pend) <= 4:
                    break
                rec['ev'].synchronize()
            pend.pop(0)
            vals = rec['host'].tolist()
            for lst, v in zip(rec['lists'], vals):
                lst.append(v)
            for key, lst, step in rec['final']:
                self._log_dict[key].append((step, np.mean(lst)))

    @property
    def log_dict(self):
        self._resolve_logs(block=True)
        return self._log_dict

    @log_dict.setter
    def log_dict(self, value):
        self.__dict__.get('_pending_logs', [])[:] = []
        self._log_dict = value

    def feed_data(self, data, need_GT=True, **kwargs):
        self.var_L = self._to_device(data['LR'])
        cur_Z = None
        if self.latent_input is not None:
            hr_size = [self.Z_size_factor * v for v in list(self.var_L.size()[2:])]
            if 'Z' in data.keys():
                cur_Z = data['Z']
            else:
                if self.cri_latent is not None:    # spatially uniform codes while the latent-control loss trains (:251-253)
                    cur_Z = torch.rand([self.var_L.size(0), self.num_latent_channels, 1, 1])
                else:
                    cur_Z = torch.rand([self.var_L.size(0), self.num_latent_channels] + hr_size).type(self.var_L.type())
                if self.opt['network_G']['latent_channels'] in ['SVD_structure_tensor', 'SVDinNormedOut_structure_tensor']:
                    cur_Z[:, -1, ...] = 2 * np.pi * cur_Z[:, -1, ...]   # (lambda0, lambda1, theta) -> structure tensor values (:256-259)
                    self.SVD = {'theta': cur_Z[:, -1, ...], 'lambda0_ratio': 1 * cur_Z[:, 0, ...], 'lambda1_ratio': 1 * cur_Z[:, 1, ...]}
                    cur_Z = SVD_2_LatentZ(cur_Z).detach()
                else:
                    cur_Z = 2 * cur_Z - 1
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
            if self.is_train and getattr(self, 'add_quantization_noise', False):
                # keeps the critic from telling real from generated images by their 8-bit quantisation (:272-273)
                data['HR'] += (torch.rand_like(data['HR']) - 0.5) / 255
            self.var_H = self._to_device(data['HR'])
            self.var_ref = self._to_device(data['ref']) if 'ref' in data else self.var_H

    @staticmethod
    def _batch_mean(t):
        """mean over the GLOBAL batch: the reference takes it after nn.DataParallel gathered the critic's outputs
        (SRRaGAN_model.py:353-354,475-476); with one process per GPU that is a 2-scalar all-reduce"""
        if parallel.world() == 1:
            return torch.mean(t)
        return parallel.global_mean_autograd(t)

    def _relativistic_terms(self, pred_a, pred_b, target_a, target_b):
        """(cri_gan(pred_a - mean(pred_b), target_a), cri_gan(pred_b - mean(pred_a), target_b)), means over the global batch
        (models/SRRaGAN_model.py:353-354,475-476).  The vanilla (BCE-with-logits) loss on CUDA logits is one fused kernel each way."""
        if self.cri_gan.gan_type == 'vanilla' and pred_a.is_cuda and pred_a.shape == pred_b.shape:
            return relativistic_bce(pred_a, pred_b, self.cri_gan.real_label_val if target_a else self.cri_gan.fake_label_val,
                                    self.cri_gan.real_label_val if target_b else self.cri_gan.fake_label_val)
        return (self.cri_gan(pred_a - self._batch_mean(pred_b), target_a), self.cri_gan(pred_b - self._batch_mean(pred_a), target_b))

    def optimize_parameters(self):
        """models/SRRaGAN_model.py:280-519: forward through CEM(G) and crop the invalid margins; discriminator step
        (critic on the real patch and on the detached fake one, relativistic average loss, Adam on the last accumulation step);
        generator step (pixel + feature + range + GAN terms scaled by the accumulation count, one backward through the critic,
        the VGG extractor, the CEM projection and the generator: dgrad + wgrad launches behind single autograd nodes)."""
        if not self.is_train:
            raise NotImplementedError('optimize_parameters needs a model built with is_train=True (no optimizer / losses exist)')
        self._resolve_logs(block=False)
        self.gradient_step_num = self.step // self.max_accumulation_steps
        first_acc = self.step % self.grad_accumulation_steps_G == 0
        last_acc = self.step % self.grad_accumulation_steps_G == (self.grad_accumulation_steps_G - 1)
        first_acc_D = self.step % self.grad_accumulation_steps_D == 0
        last_acc_D = self.step % self.grad_accumulation_steps_D == (self.grad_accumulation_steps_D - 1)
        acc_G, acc_D = self.grad_accumulation_steps_G, self.grad_accumulation_steps_D
        if self.D_exists:
            if first_acc:
                self.generator_step = self.gradient_step_num > self.D_init_iters
                if self.generator_step:
                    self.generator_step = self.gradient_step_num % max([1, self.global_D_update_ratio]) == 0
                    # when D's batch is larger than G's, G steps on the last D accumulation steps only (:292-293)
                    self.generator_step = self.generator_step and self.step % acc_D >= acc_D - acc_G
            if first_acc_D:
                self.discriminator_step = self.gradient_step_num >= -self.D_init_iters
                if self.discriminator_step:
                    if not self.verified_D_saved:
                        self.discriminator_step = True
                    else:
                        self.discriminator_step = self.gradient_step_num % max([1, np.ceil(1 / self.global_D_update_ratio)]) == 0
        # G forward: its graph is only kept when a generator step follows
        self.Set_Require_Grad_Status(self.netG, bool(not self.D_exists or self.generator_step))
        # with the optimised-Z reference loss (L_map) every batch is used twice once the generator has started learning (:314-330):
        # first with Z optimised towards the ground truth by Z_optimizer, then with the Z that was fed
        dual_steps = int(self.optimalZ_loss_type is not None and self.generator_started_learning) + 1
        for dual_num in range(dual_steps):
            optimized_Z_step = dual_num == (dual_steps - 2)
            first_dual, last_dual = dual_num == 0, dual_num == (dual_steps - 1)
            if self.CEM_net is not None and first_dual:
                self.var_H, self.var_ref = self.CEM_net.HR_unpadder(self.var_H), self.CEM_net.HR_unpadder(self.var_ref)
            if first_dual:
                static_Z = self.GetLatent() if self.latent_input is not None else None
            if optimized_Z_step:
                self.Z_optimizer.feed_data({'LR': self.var_L, 'desired': self.var_H})
                self.Z_optimizer.optimize()          # leaves self.fake_H = G(LR, optimised Z) with its graph
            else:
                self.Prepare_Input(LR_image=self.var_L, latent_input=static_Z)
                if self.D_exists and not self.generator_step:
                    with torch.no_grad():
                        self.fake_H = self.netG(self.model_input)
                else:
                    self.fake_H = self.netG(self.model_input)
            if self.CEM_net is not None:
                self.fake_H = self.CEM_net.HR_unpadder(self.fake_H)
            if not self.D_exists:
                self.generator_step = self.gradient_step_num > 0   # one idle iteration first, to save the initial validation results
            elif self.discriminator_step:
                # ---- D step (:340-414)
                self.Set_Require_Grad_Status(self.netD, True)
                if first_acc_D and first_dual:
                    self.optimizer_D.zero_grad()
                    self.l_d_real_grad_step, self.l_d_fake_grad_step, self.D_real_grad_step, self.D_fake_grad_step = [], [], [], []
                    self.D_logits_diff_grad_step = []
                if first_dual:
                    pred_d_real = self.netD(self.var_ref)
                pred_d_fake = self.netD(self.fake_H.detach())   # detach to avoid BP to G
                if self.relativistic_D:
                    assert self.opt['train']['hinge_threshold'] is None, 'Unsupported yet, should think whether it reuires special adaptation of hinge loss'
                    l_d_real, l_d_fake = self._relativistic_terms(pred_d_real, pred_d_fake, True, False)
                else:   # (x2: consistent with the SRGAN code, where the two losses are summed, :357-358)
                    if first_dual:
                        l_d_real = 2 * self.cri_gan(pred_d_real, True, self.opt['train']['hinge_threshold'])
                    l_d_fake = 2 * self.cri_gan(pred_d_fake, False, self.opt['train']['hinge_threshold'])
                l_d_total = (l_d_real + l_d_fake) / 2
                if self.opt['train']['gan_type'] == 'wgan-gp':      # (:362-371) penalty at random interpolates of real and generated
                    random_pt = torch.rand(self.var_ref.size(0), 1, 1, 1, device=self.var_ref.device)
                    interp = random_pt * self.fake_H.detach() + (1 - random_pt) * self.var_ref
                    interp.requires_grad = True
                    l_d_gp = self.l_gp_w * self.cri_gp(interp, self.netD(interp))
                    l_d_total = l_d_total + l_d_gp
                l_d_total = l_d_total / (acc_D * dual_steps)
                # logged statistics: ONE device->host read per D step, over the GLOBAL batch (the reference computes them on the
                # outputs nn.DataParallel gathered, :372-381); they gate D verification and the lr roll-back, so every rank must
                # see the same numbers
                local = torch.cat([torch.stack([l_d_real.detach(), l_d_fake.detach(), torch.mean(pred_d_real.detach()),
                                                torch.mean(pred_d_fake.detach())]).float().reshape(-1),
                                   torch.mean(pred_d_real.detach() - pred_d_fake.detach(), dim=1).float().reshape(-1)])
                allv = parallel.all_gather_cat(local.unsqueeze(0)).cpu().numpy()      # [ranks, 4 + local batch]
                stats = [float(v) for v in allv[:, :4].astype(np.float64).mean(0)] + [float(v) for v in allv[:, 4:].reshape(-1)]
                self.l_d_real_grad_step.append(stats[0])
                self.l_d_fake_grad_step.append(stats[1])
                self.D_real_grad_step.append(stats[2])
                self.D_fake_grad_step.append(stats[3])
                self.D_logits_diff_grad_step.append(list(np.asarray(stats[4:], dtype=np.float32)))
                if first_acc_D and first_dual and self.generator_step:      # D verification (:377-393): may call this generator step off
                    tr = self.opt['train']
                    if self.D_verification == 'past' and tr['D_valid_Steps_4_G_update'] > 0:
                        k = tr['D_valid_Steps_4_G_update']
                        self.generator_step = len(self.log_dict['D_logits_diff']) >= k and \
                            all([v[1] > np.log(tr['min_D_prob_ratio_4_G']) for v in self.log_dict['D_logits_diff'][-k:]]) and \
                            all([v[1] > tr['min_mean_D_correct'] for v in self.log_dict['Correctly_distinguished'][-k:]])
                    elif self.D_verification == 'convergence':
                        if not self.D_converged and self.gradient_step_num >= tr['steps_4_D_convergence']:
                            std, slope = 0, 0
                            for key in ['l_d_real', 'l_d_fake']:
                                vals = [v[1] for v in self.log_dict[key] if v[0] >= self.gradient_step_num - tr['steps_4_loss_std']]
                                [cur_slope, _], [[cur_var, _], _] = np.polyfit([i for i in range(len(vals))], vals, 1, cov=True)
                                std += 0.5 * np.sqrt(cur_var)
                                slope += 0.5 * cur_slope
                            self.D_converged = -tr['lr_change_ratio'] * np.minimum(-1e-5, slope) < std
                        self.generator_step = 1 * self.D_converged
                if self.D_verification == 'current' and self.generator_step:
                    self.generator_step = all([v > 0 for v in self.D_logits_diff_grad_step[-1]]) \
                        and np.mean(self.D_logits_diff_grad_step[-1]) > np.log(self.opt['train']['min_D_prob_ratio_4_G'])
                self.generator_step = bool(self.generator_step)     # (numpy scalars above; under numpy 2 the reference's 'current' mode trips on exactly that)
                if not self.generator_step:      # the generator's graph is not needed after all
                    self.fake_H = self.fake_H.detach()
                l_d_total.backward(retain_graph=not last_dual)      # the real batch's critic graph is shared by both dual steps
                if last_acc_D and last_dual:
                    parallel.average_gradients(self.netD.parameters(), optimizer=self.optimizer_D)
                    self.optimizer_D.step()
                    self.log_dict['l_d_real'].append((self.gradient_step_num, np.mean(self.l_d_real_grad_step)))
                    self.log_dict['l_d_fake'].append((self.gradient_step_num, np.mean(self.l_d_fake_grad_step)))
                    self.log_dict['l_d_real_fake'].append((self.gradient_step_num, np.mean(self.l_d_fake_grad_step) + np.mean(self.l_d_real_grad_step)))
                    if self.opt['train']['gan_type'] == 'wgan-gp':
                        self.log_dict['l_d_gp'].append((self.gradient_step_num, l_d_gp.item()))
                    self.log_dict['D_real'].append((self.gradient_step_num, np.mean(self.D_real_grad_step)))
                    self.log_dict['D_fake'].append((self.gradient_step_num, np.mean(self.D_fake_grad_step)))
                    self.log_dict['D_logits_diff'].append((self.gradient_step_num, np.mean(self.D_logits_diff_grad_step)))
                    self.log_dict['Correctly_distinguished'].append((self.gradient_step_num, np.mean([v0 > 0 for v1 in self.D_logits_diff_grad_step for v0 in v1])))
            if self.generator_step:
                # ---- G step (:417-519)
                self.generator_started_learning = True
                if self.D_exists:
                    self.Set_Require_Grad_Status(self.netD, False)
                self.Set_Require_Grad_Status(self.netG, True)
                if first_acc and first_dual:
                    self.optimizer_G.zero_grad()
                    self.l_g_pix_grad_step, self.l_g_range_grad_step, self.l_g_fea_grad_step, self.l_g_gan_grad_step = [], [], [], []
                    self.l_g_latent_grad_step, self.l_g_optimalZ_grad_step = [], []
                l_g_total = 0
                if self.cri_pix:
                    l_g_pix = self.cri_pix(self.fake_H, self.var_H)
                    l_g_total = l_g_total + self.l_pix_w * l_g_pix / (acc_G * dual_steps)
                if self.cri_fea:   # perceptual loss: VGG features of the real image (no graph) and of the generated one
                    real_fea = self.netF(self.var_H).detach()
                    fake_fea = self.netF(self.fake_H)
                    l_g_fea = self.cri_fea(fake_fea, real_fea)
                    l_g_total = l_g_total + self.l_fea_w * l_g_fea / (acc_G * dual_steps)
                if self.cri_range:
                    l_g_range = self.cri_range(self.fake_H)
                    l_g_total = l_g_total + self.l_range_w * l_g_range / (acc_G * dual_steps)
                if self.cri_latent and last_dual:   # latent-control loss, on the pass with the Z that was fed (:455-461)
                    l_g_latent = self.cri_latent({'SR': self.fake_H, 'HR': self.var_H, 'Z': static_Z}).mean(0)
                    l_g_total = l_g_total + self.l_latent_w * l_g_latent.mean() / acc_G
                    self.l_g_latent_grad_step.append([v.item() for v in l_g_latent])
                if self.cri_optimalZ and first_dual:   # L_map: with the optimised Z the output should reach the ground truth (:462-465)
                    l_g_optimalZ = self.cri_optimalZ(self.fake_H, self.var_H)
                    l_g_total = l_g_total + self.l_g_optimalZ_w * l_g_optimalZ / acc_G
                    self.l_g_optimalZ_grad_step.append(l_g_optimalZ.item())
                if self.D_exists:   # G gan loss (:466-479)
                    pred_g_fake = self.netD(self.fake_H)
                    if self.relativistic_D:
                        # (the reference re-uses the D step's variable here, :474: with two dual steps the second D step therefore sees the
                        #  real batch's logits DETACHED - mirrored, it changes the critic's gradient)
                        pred_d_real = self.netD(self.var_ref).detach()
                        l_g_real, l_g_fake = self._relativistic_terms(pred_d_real, pred_g_fake, False, True)
                        l_g_gan = self.l_gan_w * (l_g_real + l_g_fake) / 2 / (acc_G * dual_steps)
                    else:
                        l_g_gan = self.l_gan_w * self.cri_gan(pred_g_fake, True) / (acc_G * dual_steps)
                    l_g_total = l_g_total + l_g_gan
                l_g_total.backward()
                # logged scalars of the G step: one device->host read instead of one .item() sync per term
                terms = [(self.cri_fea, 'l_g_fea_grad_step', l_g_fea if self.cri_fea else None), (self.cri_pix, 'l_g_pix_grad_step', l_g_pix if self.cri_pix else None),
                         (self.cri_gan, 'l_g_gan_grad_step', l_g_gan if self.cri_gan else None), (self.cri_range, 'l_g_range_grad_step', l_g_range if self.cri_range else None)]
                terms = [(name, v) for on, name, v in terms if on]
                deferred = None
                if terms:
                    stacked = torch.stack([v.detach().float().reshape(()) for _, v in terms])
                    if stacked.is_cuda and os.environ.get('ESR_SYNC_LOGS', '0') != '1':
                        # one small device->host copy per step, read when it has arrived (or when someone looks at the logs)
                        host = torch.empty(len(terms), dtype=torch.float32, pin_memory=True)
                        host.copy_(stacked, non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record()
                        deferred = {'ev': ev, 'host': host, 'lists': [getattr(self, name) for name, _ in terms], 'final': []}
                        self.__dict__.setdefault('_pending_logs', []).append(deferred)
                    else:
                        for (name, _), v in zip(terms, stacked.cpu().tolist()):
                            getattr(self, name).append(v)
                if last_acc and last_dual:
                    parallel.average_gradients([p for p in self.netG.parameters() if p.requires_grad], optimizer=self.optimizer_G)
                    self.optimizer_G.step()
                    self.generator_changed = True
                    for on, key, lst in ((self.cri_pix, 'l_g_pix', self.l_g_pix_grad_step), (self.cri_fea, 'l_g_fea', self.l_g_fea_grad_step),
                                         (self.cri_range, 'l_g_range', self.l_g_range_grad_step), (self.cri_gan, 'l_g_gan', self.l_g_gan_grad_step)):
                        if not on:
                            continue
                        if deferred is not None:
                            deferred['final'].append((key, lst, self.gradient_step_num))
                        else:
                            self.log_dict[key].append((self.gradient_step_num, np.mean(lst)))
                    if self.cri_latent:
                        for ch in range(self.num_latent_channels):
                            self.log_dict['l_g_latent_%d' % ch].append((self.gradient_step_num, np.mean([v[ch] for v in self.l_g_latent_grad_step])))
                    if self.cri_optimalZ:
                        self.log_dict['l_g_optimalZ'].append((self.gradient_step_num, np.mean(self.l_g_optimalZ_grad_step)))
        self.step += 1

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
        out = OrderedDict()
        for k, v in self.log_dict.items():
            if len(v) > 0:
                out[k] = v[-1][1] if (isinstance(v[-1], tuple) or len(v[-1]) > 1) else v[-1]
        return out

    # ---- what the reference's train.py calls around the step (validation, logs, loss-driven lr drop) ------------------------
    def perform_validation(self, data_loader, cur_Z, print_rlt, first_eval, save_images):
        """models/SRRaGAN_model.py:533-590: run every validation image through `test()`, PSNR against HR on [0,255] images, a collage of
        centre crops per call (and of the HR images on the first one).  Image conversion / PSNR / PNG writing follow utils/util.py:196-232,340-347."""
        psnrs, collage, gt_collage, sr_images = [], [], [], []
        idx = 0
        if save_images:
            num = len(data_loader.dataset)
            rows = int(np.floor(np.sqrt(num)))
            while rows > 1 and np.round(num / rows) != num / rows:
                rows -= 1
            patch = min([min(im['HR'].shape[1:]) for im in data_loader.dataset]) - 2
        for val_data in data_loader:
            if save_images and idx % rows == 0:
                collage.append([])
                gt_collage.append([])
            idx += 1
            val_data['Z'] = cur_Z
            self.feed_data(val_data)
            self.test()
            visuals = self.get_current_visuals()
            sr_img, gt_img = 255 * _tensor2img(visuals['SR']), 255 * _tensor2img(visuals['HR'])
            sr_images.append(sr_img)
            psnrs.append(_psnr(sr_img, gt_img))
            if save_images:
                m = ((np.array(sr_img.shape[:2]) - patch) / 2).astype(np.int32)
                collage[-1].append(np.clip(sr_img[m[0]:-m[0], m[1]:-m[1], ...], 0, 255).astype(np.uint8))
                if first_eval:
                    gt_collage[-1].append(np.clip(gt_img[m[0]:-m[0], m[1]:-m[1], ...], 0, 255).astype(np.uint8))
        avg_psnr = 1 * np.mean(psnrs)
        if save_images:
            self.generator_changed = False
            if 'im_collages' not in self.__dict__:
                self.im_collages = []
            self.im_collages.append(np.concatenate([np.concatenate(col, 0) for col in collage], 1))
            name = '{:d}_{}PSNR{:.3f}.png'.format(self.gradient_step_num, ('Z' + str(cur_Z)) if self.opt['network_G']['latent_input'] else '', avg_psnr)
            _save_png(self.im_collages[-1], os.path.join(self.opt['path']['val_images'], name))
            if first_eval:
                _save_png(np.concatenate([np.concatenate(col, 0) for col in gt_collage], 1), os.path.join(self.opt['path']['val_images'], 'GT_HR.png'))
        print_rlt['psnr'] += avg_psnr
        return sr_images

    def update_learning_rate(self, cur_step=None):
        """models/SRRaGAN_model.py:592-632 (returns "learning rate too low"): once enough discriminator steps are logged, the standard
        deviation of the D loss over the last `steps_4_loss_std` steps is recorded; above `std_4_lr_drop` the run rolls back to the
        checkpoint before that window and every optimizer's lr is multiplied by `lr_gamma`."""
        tr = self.opt['train']
        if not self.D_exists or tr['steps_4_loss_std'] is None:
            return False
        n = tr['steps_4_loss_std']
        if len(self.log_dict['D_logits_diff']) >= n:
            vals = [(v[1] + self.log_dict['l_d_fake'][i][1]) / 2 for i, v in enumerate(self.log_dict['l_d_real']) if v[0] >= cur_step - n]
            self.log_dict.setdefault('D_loss_STD', []).append([self.gradient_step_num, np.std(vals)])
            reduce_lr = (tr['std_4_lr_drop'] is not None) and self.log_dict['D_loss_STD'][-1][1] > tr['std_4_lr_drop']
        else:
            reduce_lr = False
        if len(self.log_dict['D_logits_diff']) < 2 * n or self.log_dict['D_logits_diff'][0][0] > cur_step - n:   # not before a minimal number of steps
            return False
        if reduce_lr:
            cur_LR = [o.param_groups[0]['lr'] for o in self.optimizers]
            self.load(max_step=cur_step - n, resume_train=True)
            for k, optimizer in enumerate(self.optimizers):
                for group in optimizer.param_groups:
                    group['lr'] = cur_LR[k] * tr['lr_gamma']
                    if group['lr'] < 1e-8:
                        return True
            lrs = {'lr_G': self.optimizer_G.param_groups[0]['lr'], 'lr_D': self.optimizer_D.param_groups[0]['lr']}
            print('LR(D) reduced to %.2e, LR(G) reduced to %.2e.' % (lrs['lr_D'], lrs['lr_G']))
            np.savez(os.path.join(self.log_path, 'lr.npz'), step_num=cur_step, **lrs)
            self.log_dict['LR_decrease'].append([self.step // self.max_accumulation_steps, lrs])
        return False

    def save_log(self):
        """logs.npz in the reference's layout (:644-651): every log series plus D_verified / verified_D_saved / lr_G / lr_D"""
        blob = dict(self.log_dict)
        for attr in ('D_verified', 'verified_D_saved', 'lr_G', 'lr_D'):
            if attr in self.__dict__:
                blob[attr] = getattr(self, attr)
        np.savez(os.path.join(self.log_path, 'logs.npz'), **{k: np.array(v, dtype=object) if k == 'LR_decrease' else v for k, v in blob.items()})
        if self.cri_latent is not None and hasattr(self.cri_latent, 'collected_ratios'):
            np.savez(os.path.join(self.log_path, 'collected_stats.npz'), *self.cri_latent.collected_ratios)

    def load_log(self, max_step=None):
        """(:653-675) the inverse of save_log, optionally truncated to gradient steps <= max_step (roll-back after an lr drop)"""
        from collections import deque
        loaded = np.load(os.path.join(self.log_path, 'logs.npz'), allow_pickle=True)
        self.log_dict = OrderedDict((k, []) for k in self.log_dict.keys())
        for key in loaded.files:
            if key in ('D_verified', 'verified_D_saved', 'lr_G', 'lr_D'):
                # np.load hands back 0-d arrays: cast (reference: bool() at :209), or they end up pickled inside the optimizers'
                # param_groups and the next checkpoint cannot be read back (torch.load weights_only)
                setattr(self, key, bool(loaded[key]) if key in ('D_verified', 'verified_D_saved') else float(loaded[key]))
                continue
            self.log_dict[key] = [tuple(v) for v in loaded[key]] if key == 'psnr_val' else list(loaded[key])
            if max_step is not None:
                self.log_dict[key] = [pair for pair in self.log_dict[key] if pair[0] <= max_step]
        if self.cri_latent is not None and hasattr(self.cri_latent, 'collected_ratios'):
            stats = np.load(os.path.join(self.log_path, 'collected_stats.npz'))
            for i, f in enumerate(stats.files):
                self.cri_latent.collected_ratios[i] = deque(stats[f], maxlen=self.cri_latent.collected_ratios[i].maxlen)

    def display_log_figure(self):
        """one PDF per logged series (models/base_model.py:211-274), written when matplotlib is installed; plotting is not part of the
        accelerated path, a missing matplotlib only skips the figures"""
        try:
            import matplotlib
            matplotlib.use('Agg')
            import matplotlib.pyplot as plt
        except Exception:
            return
        for key, series in self.log_dict.items():
            if key == 'LR_decrease' or len(series) == 0:
                continue
            steps, vals = np.array([v[0] for v in series]), np.array([float(v[1]) for v in series])
            plt.figure(1)
            plt.clf()
            plt.plot(steps, vals)
            for decrease in self.log_dict.get('LR_decrease', []):
                plt.plot([decrease[0], decrease[0]], [vals.min(), vals.max()], 'k')
            plt.xlabel('Steps')
            plt.legend([key + ' (%.2e)' % vals.mean()], loc='best')
            plt.savefig(os.path.join(self.log_path, 'logs_%s.pdf' % key))

    # ---- checkpoints --------------------------------------------------------------------------------
    def load(self, max_step=None, resume_train=None):
        """models/SRRaGAN_model.py:732-771.  Self-trained checkpoints (`<step>_G.pth` [+ `<step>_D.pth`] in path.models) are taken
        when resuming training (`train.resume` / resume_train: optimizer states and the step counter come back too), when a
        `max_step` is asked for, or when testing; otherwise the pretrained paths.  (With nothing to resume from, the reference
        fails on an empty list; here the pretrained path is used.)"""
        path = self.opt['path']
        models_dir = path['models'] if path is not None else None
        resume_training = resume_train if resume_train is not None else bool(self.is_train and self.opt['train']['resume'])
        own = [n for n in os.listdir(models_dir) if '_G.pth' in n] if (models_dir and os.path.isdir(models_dir)) else []
        step_of = lambda n: int(re.search(r'(\d)+(?=_G.pth)', n).group(0))
        own = sorted(own, key=step_of)
        if max_step is not None:
            own = [n for n in own if step_of(n) <= max_step]
        load_own = (max_step is not None or resume_training or not self.is_train) and len(own) > 0
        if load_own:
            name = own[-1]
            loaded_step = step_of(name)
            d_path = os.path.join(models_dir, '%d_D.pth' % loaded_step)
            if self.is_train:
                self.step = (loaded_step + 1) * self.max_accumulation_steps
                print('Resuming training with model for G [{:s}] ...'.format(os.path.join(models_dir, name)))
                self.load_network(os.path.join(models_dir, name), self.netG, optimizer=self.optimizer_G)
                if self.log_path is not None and os.path.exists(os.path.join(self.log_path, 'logs.npz')):
                    self.load_log(max_step=loaded_step)       # the logs roll back with the weights (:746)
                if self.D_exists:
                    print('Resuming training with model for D [{:s}] ...'.format(d_path))
                    self.load_network(d_path, self.netD, optimizer=self.optimizer_D)
            else:
                print('Testing model for G [{:s}] ...'.format(os.path.join(models_dir, name)))
                self.load_network(os.path.join(models_dir, name), self.netG)
                if 'netD' in self.__dict__ and os.path.exists(d_path):   # when running from the GUI
                    print('Loading also model for D [{:s}] ...'.format(d_path))
                    self.load_network(d_path, self.netD)
                self.gradient_step_num = loaded_step
            return
        if path is not None and path['pretrained_model_G'] is not None:
            print('loading model for G [{:s}] ...'.format(path['pretrained_model_G']))
            self.load_network(path['pretrained_model_G'], self.netG)
        if self.is_train and self.D_exists and path is not None and path['pretrained_model_D'] is not None:
            print('loading model for D [{:s}] ...'.format(path['pretrained_model_D']))
            self.load_network(path['pretrained_model_D'], self.netD, optimizer=self.optimizer_D)

    def save(self, iter_label):
        saving_path = self.save_network(self.save_dir, self.netG, 'G', iter_label, self.optimizer_G)
        if self.D_exists:
            self.save_network(self.save_dir, self.netD, 'D', iter_label, self.optimizer_D)
        return saving_path
