"""SRModel — the G-only model that options/test/test_sr.json ("model": "sr") instantiates
(reference: codes/SRN/models/SR_model.py:19-173).  netG runs on the dasr_b200 kernels."""
import logging
from collections import OrderedDict

import torch
import torch.nn as nn
from torch.optim import lr_scheduler

from dasr_b200.srn.utils.util import forward_chop
from . import networks
from .base_model import BaseModel
from .modules import loss as L

logger = logging.getLogger('base')


class SRModel(BaseModel):
    def __init__(self, opt):
        super().__init__(opt)
        train_opt = opt['train']
        self.chop = opt['chop']
        self.scale = opt['scale']
        self.val_lpips = opt['val_lpips']
        self.netG = networks.define_G(opt).to(self.device)
        self.load()
        if self.is_train:
            self.netG.train()
            loss_type = train_opt['pixel_criterion']
            if loss_type == 'l1':
                self.cri_pix = L.L1Loss().to(self.device)
            elif loss_type == 'l2':
                self.cri_pix = L.MSELoss().to(self.device)
            else:
                raise NotImplementedError('Loss type [{:s}] is not recognized.'.format(loss_type))
            self.l_pix_w = train_opt['pixel_weight']
            wd_G = train_opt['weight_decay_G'] if train_opt['weight_decay_G'] else 0
            optim_params = []
            for k, v in self.netG.named_parameters():
                if v.requires_grad:
                    optim_params.append(v)
                else:
                    logger.warning('Params [{:s}] will not optimize.'.format(k))
            self.optimizer_G = torch.optim.Adam(optim_params, lr=train_opt['lr_G'], weight_decay=wd_G)
            self.optimizers.append(self.optimizer_G)
            if train_opt['lr_scheme'] == 'MultiStepLR':
                for optimizer in self.optimizers:
                    self.schedulers.append(lr_scheduler.MultiStepLR(optimizer, train_opt['lr_steps'], train_opt['lr_gamma']))
            else:
                raise NotImplementedError('MultiStepLR learning rate scheme is enough.')
            self.log_dict = OrderedDict()
        self.print_network()
        if self.val_lpips:
            # LPIPS (AlexNet trunk) is outside the hot path and its weights are not available offline
            logger.warning('val_lpips requested: LPIPS is not part of the B200 path; LPIPS is reported as nan')

    def feed_data(self, data, need_HR=True):
        self.var_L = data['LR'].to(self.device)
        if 'HR' in data:
            self.real_H = data['HR'].to(self.device)

    def optimize_parameters(self, step):
        self.optimizer_G.zero_grad()
        self.fake_H = self.netG(self.var_L)
        l_pix = self.l_pix_w * self.cri_pix(self.fake_H, self.real_H)
        l_pix.backward()
        self.optimizer_G.step()
        self.log_dict['l_pix'] = l_pix.item()

    def test(self):
        self.netG.eval()
        with torch.no_grad():
            if self.chop:
                self.fake_H = forward_chop(self.var_L, self.scale, self.netG)
            else:
                self.fake_H = self.netG(self.var_L)
            if self.val_lpips:
                self.LPIPS = torch.tensor(float('nan'))
        self.netG.train()

    def test_x8(self):
        """x8 self-ensemble (flips + transpose), averaged."""
        self.netG.eval()

        def tf(v, op):
            if op == 'v':
                return v.flip(3)
            if op == 'h':
                return v.flip(2)
            return v.transpose(2, 3)

        with torch.no_grad():
            lr_list = [self.var_L]
            for op in 'v', 'h', 't':
                lr_list.extend([tf(t, op).contiguous() for t in lr_list])
            sr_list = [self.netG(a) for a in lr_list]
            for i in range(len(sr_list)):
                if i > 3:
                    sr_list[i] = tf(sr_list[i], 't')
                if i % 4 > 1:
                    sr_list[i] = tf(sr_list[i], 'h')
                if (i % 4) % 2 == 1:
                    sr_list[i] = tf(sr_list[i], 'v')
            self.fake_H = torch.cat(sr_list, dim=0).mean(dim=0, keepdim=True)
        self.netG.train()

    def get_current_log(self):
        return self.log_dict

    def get_current_visuals(self, need_HR=True):
        out = OrderedDict()
        out['LR'] = self.var_L.detach()[0].float().cpu()
        out['SR'] = self.fake_H.detach()[0].float().cpu()
        if need_HR:
            out['HR'] = self.real_H.detach()[0].float().cpu()
        if self.val_lpips:
            out['LPIPS'] = self.LPIPS.detach().float().cpu()
        return out

    def print_network(self):
        s, n = self.get_network_description(self.netG)
        if isinstance(self.netG, nn.DataParallel):
            name = '{} - {}'.format(self.netG.__class__.__name__, self.netG.module.__class__.__name__)
        else:
            name = '{}'.format(self.netG.__class__.__name__)
        logger.info('Network G structure: {}, with parameters: {:,d}'.format(name, n))
        logger.info(s)

    def load(self):
        load_path_G = self.opt['path']['pretrain_model_G']
        if load_path_G is not None:
            logger.info('Loading pretrained model for G [{:s}] ...'.format(load_path_G))
            self.load_network(load_path_G, self.netG)

    def save(self, iter_step):
        self.save_network(self.netG, 'G', iter_step)
