"""SRModel — the generator-only model behind options/test/test_sr.json ("model": "sr"); reference: codes/SRN/models/
SR_model.py:19-173.  Same attributes and methods; netG runs on the dasr_b200 kernels."""
import logging
from collections import OrderedDict

import torch

from dasr_b200.srn.utils.util import forward_chop
from . import networks
from .base_model import BaseModel

logger = logging.getLogger('base')

# the eight dihedral views of test_x8: (transpose?, flip H?, flip W?) — index bits as in the reference's loop order
_X8 = [(t, h, v) for t in (False, True) for h in (False, True) for v in (False, True)]


def _view(img, t, h, v, inverse=False):
    """Apply (or undo) one dihedral transform on an NCHW batch."""
    steps = [('v', v), ('h', h), ('t', t)]
    if inverse:
        steps.reverse()
    for kind, on in steps:
        if on:
            img = img.flip(3) if kind == 'v' else img.flip(2) if kind == 'h' else img.transpose(2, 3)
    return img.contiguous()


def _val_lpips(metric, a, b, device):
    """LPIPS of the 8-bit round trip of two image batches (SR_model.py:95-99 / DASR_model.py:340-344): tensor2img
    (clamp, x255, round, BGR) -> RGB -> im2tensor ([-1, 1]) -> net-lin alex distance of the first image."""
    from dasr_b200.lpips import im2tensor
    from dasr_b200.srn.utils import util
    ia, ib = util.tensor2img(a.detach().float().cpu().clone()), util.tensor2img(b.detach().float().cpu().clone())
    ia, ib = ia[:, :, [2, 1, 0]], ib[:, :, [2, 1, 0]]
    return metric(im2tensor(ia).to(device), im2tensor(ib).to(device))[0][0][0][0]


class SRModel(BaseModel):
    def __init__(self, opt):
        super().__init__(opt)
        self.chop, self.scale, self.val_lpips = opt['chop'], opt['scale'], opt['val_lpips']
        self.netG = networks.define_G(opt).to(self.device)
        self.load()
        if self.is_train:
            self._init_training(opt['train'])
        self.print_network()
        if self.val_lpips:      # SR_model.py:66-67: PerceptualLoss(model='net-lin', net='alex')
            from dasr_b200.lpips import PerceptualLoss
            self.cri_fea_lpips = PerceptualLoss(model='net-lin', net='alex').to(self.device)

    def _init_training(self, cfg):
        self.netG.train()
        self.cri_pix = self._criterion(cfg['pixel_criterion'])
        self.l_pix_w = cfg['pixel_weight']
        self.optimizer_G = self._adam(self.netG, cfg['lr_G'], cfg['weight_decay_G'])
        self._make_schedulers(cfg)
        self.log_dict = OrderedDict()

    def feed_data(self, data, need_HR=True):
        self.var_L = self._to_device(data['LR'])
        if 'HR' in data:
            self.real_H = self._to_device(data['HR'])

    def optimize_parameters(self, step):
        self.optimizer_G.zero_grad()
        self.fake_H = self.netG(self.var_L)
        loss = self.l_pix_w * self.cri_pix(self.fake_H, self.real_H)
        loss.backward()
        self.optimizer_G.step()
        self.log_dict['l_pix'] = loss.item()

    def test(self):
        self._log_eval_precision(self.netG)
        self.netG.eval()
        with torch.no_grad():
            self.fake_H = forward_chop(self.var_L, self.scale, self.netG) if self.chop else self.netG(self.var_L)
            if self.val_lpips:
                self.LPIPS = _val_lpips(self.cri_fea_lpips, self.real_H, self.fake_H, self.device)
        self.netG.train()

    def test_x8(self):
        """x8 self-ensemble: the generator on the 8 dihedral views of the input, mapped back and averaged."""
        self.netG.eval()
        with torch.no_grad():
            outs = [_view(self.netG(_view(self.var_L, *cfg)), *cfg, inverse=True) for cfg in _X8]
            self.fake_H = torch.cat(outs, dim=0).mean(dim=0, keepdim=True)
        self.netG.train()

    def get_current_log(self):
        return self.log_dict

    def get_current_visuals(self, need_HR=True):
        first = lambda t: t.detach()[0].float().cpu()
        vis = OrderedDict(LR=first(self.var_L), SR=first(self.fake_H))
        if need_HR:
            vis['HR'] = first(self.real_H)
        if self.val_lpips:
            vis['LPIPS'] = self.LPIPS.detach().float().cpu()
        return vis

    def print_network(self):
        self._log_network(self.netG, 'G')

    def load(self):
        path = self.opt['path']['pretrain_model_G']
        if path is not None:
            logger.info('Loading pretrained model for G [{:s}] ...'.format(path))
            self.load_network(path, self.netG)

    def save(self, iter_step):
        self.save_network(self.netG, 'G', iter_step)
