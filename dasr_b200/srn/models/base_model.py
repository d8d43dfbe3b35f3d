"""BaseModel: device pick, LR stepping, network / training-state save + load
(reference: codes/SRN/models/base_model.py:6-85; file names and dict layout are the checkpoint contract)."""
import os

import torch
import torch.nn as nn


class BaseModel():
    def __init__(self, opt):
        self.opt = opt
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

    def update_learning_rate(self):
        for scheduler in self.schedulers:
            scheduler.step()

    def get_current_learning_rate(self):
        s = self.schedulers[0]
        return (s.get_last_lr() if hasattr(s, 'get_last_lr') else s.get_lr())[0]

    @staticmethod
    def _unwrap(network):
        return network.module if isinstance(network, nn.DataParallel) else network

    def get_network_description(self, network):
        network = self._unwrap(network)
        return str(network), sum(p.numel() for p in network.parameters())

    def save_network(self, network, network_label, iter_step):
        path = os.path.join(self.opt['path']['models'], '{}_{}.pth'.format(iter_step, network_label))
        sd = self._unwrap(network).state_dict()
        torch.save({k: v.cpu() for k, v in sd.items()}, path)

    def load_network(self, load_path, network, strict=True):
        self._unwrap(network).load_state_dict(torch.load(load_path, map_location='cpu'), strict=strict)

    def save_training_state(self, epoch, iter_step):
        state = {'epoch': epoch, 'iter': iter_step,
                 'schedulers': [s.state_dict() for s in self.schedulers],
                 'optimizers': [o.state_dict() for o in self.optimizers]}
        torch.save(state, os.path.join(self.opt['path']['training_state'], '{}.state'.format(iter_step)))

    def resume_training(self, resume_state, opt=None):
        ro, rs = resume_state['optimizers'], resume_state['schedulers']
        assert len(ro) == len(self.optimizers), 'Wrong lengths of optimizers'
        assert len(rs) == len(self.schedulers), 'Wrong lengths of schedulers'
        for o, st in zip(self.optimizers, ro):
            o.load_state_dict(st)
        for s, st in zip(self.schedulers, rs):
            s.load_state_dict(st)
