"""BaseModel: checkpoint format and state-dict remapping of the reference (models/base_model.py:114-190).

  * files are torch.save({'model_state_dict', 'optimizer_state_dict'}); a bare state dict is accepted as well (:132-135);
  * a checkpoint of a bare generator loads into a CEM-wrapped one (CEMnet.Adjust_State_Dict_Keys);
  * keys are matched POSITIONALLY (older checkpoints used different module names, :156-163);
  * a checkpoint without latent inputs initialises a latent model: the extra input channels are the FIRST ones of every
    conv and get zero weights (:164-175); the CEM's own filters are never loaded (:182-183)."""
import collections
import os

import numpy as np
import torch
import torch.nn as nn

import CEM.CEMnet as CEMnet


class BaseModel():
    def __init__(self, opt):
        self.opt = opt
        self.save_dir = opt['path']['models']
        self.device = torch.device('cuda' if opt['gpu_ids'] is not None else 'cpu')
        self.is_train = opt['is_train']
        self.schedulers = []
        self.optimizers = []

    def feed_data(self, data):
        pass

    def optimize_parameters(self):
        pass

    def get_current_visuals(self):
        pass

    def get_current_losses(self):
        pass

    def print_network(self):
        pass

    def save(self, label):
        pass

    def load(self):
        pass

    def update_learning_rate(self, cur_step=None):
        for scheduler in self.schedulers:
            scheduler.step(cur_step)

    def get_current_learning_rate(self):
        return self.optimizers[0].param_groups[0]['lr']

    def get_network_description(self, network):
        if isinstance(network, nn.DataParallel):
            network = network.module
        return {'s': str(network), 'n': sum(p.numel() for p in network.parameters())}

    def save_network(self, save_dir, network, network_label, iter_label, optimizer):
        save_path = os.path.join(save_dir, '{}_{}.pth'.format(iter_label, network_label))
        if isinstance(network, nn.DataParallel):
            network = network.module
        model_state_dict = collections.OrderedDict((k, v.cpu()) for k, v in network.state_dict().items())
        torch.save({'model_state_dict': model_state_dict, 'optimizer_state_dict': optimizer.state_dict() if optimizer is not None else {}},
                   save_path)
        return save_path

    def load_network(self, load_path, network, strict=False, optimizer=None):
        if isinstance(network, nn.DataParallel):
            network = network.module
        loaded = torch.load(load_path, map_location='cpu')
        if 'optimizer_state_dict' in loaded.keys():
            if optimizer is not None:
                optimizer.load_state_dict(loaded['optimizer_state_dict'])
            loaded = loaded['model_state_dict']
        if self.opt['network_G']['CEM_arch']:
            loaded = CEMnet.Adjust_State_Dict_Keys(loaded, network.state_dict())
        loaded = self.process_loaded_state_dict(loaded_state_dict=loaded, current_state_dict=network.state_dict())
        network.load_state_dict(loaded, strict=strict)

    def Set_Require_Grad_Status(self, network, status):
        for p in network.parameters():
            p.requires_grad = status

    def process_loaded_state_dict(self, loaded_state_dict, current_state_dict):
        out = collections.OrderedDict()
        current_keys = list(current_state_dict.keys())
        assert len(current_keys) == len(loaded_state_dict), 'Loaded model and current one should have the same number of parameters'
        renamed = extended = 0
        z = self.num_latent_channels if (getattr(self, 'latent_input', None) is not None) else 0
        cem_ops = self.CEM_net.OP_names if (getattr(self, 'CEM_net', None) is not None and getattr(self, 'CEM_arch', False)) else []
        for key, cur_key in zip(loaded_state_dict.keys(), current_keys):
            src, dst = loaded_state_dict[key], current_state_dict[cur_key]
            if key != cur_key:
                assert src.shape[:1] + src.shape[2:] == dst.shape[:1] + dst.shape[2:], 'Unmatching parameter sizes after changing parameter key name'
                renamed += 1
            if z > 0 and 'weight' in key and src.dim() > 1 and 0 < dst.shape[1] - src.shape[1] <= z:
                extra = dst.shape[1] - src.shape[1]
                out[cur_key] = torch.cat([torch.zeros([dst.shape[0], extra] + list(dst.shape[2:]), dtype=src.dtype), src.cpu()], 1)
                extended += 1
            elif any(op in key for op in cem_ops):
                continue
            else:
                out[cur_key] = src
        if renamed:
            print('Warning: Modified %d key names due to the change to using ModuleLists' % renamed)
        if extended:
            print('Warning: %d model weights were augmented with zeros to accommodate for larger inputs' % extended)
        return out
