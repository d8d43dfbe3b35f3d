"""BaseModel of the SRN mirror: device selection, schedulers, checkpoint files (reference: codes/SRN/models/
base_model.py:6-85 — the file names `{iter}_{label}.pth`, `{iter}.state` and their dict layout are the checkpoint
contract), plus the small factories the model classes share."""
import logging
import os

import torch
import torch.nn as nn
from torch.optim import lr_scheduler

logger = logging.getLogger('base')


def unwrap(net):
    """The module behind an optional nn.DataParallel wrapper."""
    return net.module if isinstance(net, nn.DataParallel) else net


class BaseModel():
    def __init__(self, opt):
        self.opt = opt
        self.is_train = opt['is_train']
        self.device = torch.device('cpu' if opt['gpu_ids'] is None else 'cuda')
        self.optimizers, self.schedulers = [], []

    # --- interface stubs the concrete models override -------------------------------------------------
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

    # --- shared factories ------------------------------------------------------------------------------
    def _criterion(self, kind, what='Loss'):
        """'l1' | 'l2' -> fused value+gradient loss module on this model's device."""
        from .modules import loss as L
        table = {'l1': L.L1Loss, 'l2': L.MSELoss}
        if kind not in table:
            raise NotImplementedError('{} type [{:s}] not recognized.'.format(what, str(kind)))
        return table[kind]().to(self.device)

    def _adam(self, net, lr, weight_decay, beta1=None):
        """Adam over the trainable parameters of `net` (frozen ones are reported like the reference does)."""
        trainable = []
        for name, p in net.named_parameters():
            if p.requires_grad:
                trainable.append(p)
            else:
                logger.warning('Params [{:s}] will not optimize.'.format(name))
        kw = {} if beta1 is None else {'betas': (beta1, 0.999)}
        opt = torch.optim.Adam(trainable, lr=lr, weight_decay=weight_decay or 0, **kw)
        self.optimizers.append(opt)
        return opt

    def _make_schedulers(self, train_opt):
        if train_opt['lr_scheme'] != 'MultiStepLR':
            raise NotImplementedError('MultiStepLR learning rate scheme is enough.')
        self.schedulers.extend(lr_scheduler.MultiStepLR(o, train_opt['lr_steps'], train_opt['lr_gamma']) for o in self.optimizers)

    def _log_network(self, net, tag):
        text, count = self.get_network_description(net)
        inner = unwrap(net).__class__.__name__
        name = '{} - {}'.format(net.__class__.__name__, inner) if isinstance(net, nn.DataParallel) else inner
        logger.info('Network {} structure: {}, with parameters: {:,d}'.format(tag, name, count))
        logger.info(text)

    # --- learning rate ----------------------------------------------------------------------------------
    def update_learning_rate(self):
        for sch in self.schedulers:
            sch.step()

    def get_current_learning_rate(self):
        first = self.schedulers[0]
        lrs = first.get_last_lr() if hasattr(first, 'get_last_lr') else first.get_lr()
        return lrs[0]

    # --- checkpoints --------------------------------------------------------------------------------------
    _unwrap = staticmethod(unwrap)

    def get_network_description(self, network):
        net = unwrap(network)
        return str(net), sum(p.numel() for p in net.parameters())

    def _to_device(self, t):
        """Host -> device copy of a batch tensor: asynchronous when the source is pinned (DataLoader pin_memory, bench.py) so
        the host does not stall behind the previous step's GPU work; pageable sources copy synchronously as in the reference."""
        return t.to(self.device, non_blocking=bool(getattr(t, 'is_pinned', lambda: False)()))

    def _log_eval_precision(self, net):
        """Once per model: which arithmetic produces the validation / test images (the reference validates in fp32; here
        the default is the bf16 tensor-core path, DASR_B200_PRECISION=fp16 | fp32 for closer metrics)."""
        if getattr(self, '_eval_precision_logged', False):
            return
        self._eval_precision_logged = True
        import logging
        import os
        net = net.module if hasattr(net, 'module') else net
        prec = getattr(net, 'precision', None) or os.environ.get('DASR_B200_PRECISION', 'bf16')
        logging.getLogger('base').info('dasr_b200: test() / validation forward runs in precision [%s]' % prec)

    @staticmethod
    def _is_writer():
        """Under torch.distributed every replica holds the same weights: only rank 0 writes checkpoint files."""
        import torch.distributed as dist
        return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0

    def save_network(self, network, network_label, iter_step):
        if not self._is_writer():
            return
        target = os.path.join(self.opt['path']['models'], '{}_{}.pth'.format(iter_step, network_label))
        torch.save({k: t.cpu() for k, t in unwrap(network).state_dict().items()}, target)

    def load_network(self, load_path, network, strict=True):
        unwrap(network).load_state_dict(torch.load(load_path, map_location='cpu'), strict=strict)

    def save_training_state(self, epoch, iter_step):
        if not self._is_writer():
            return
        snapshot = {'epoch': epoch, 'iter': iter_step,
                    'schedulers': [s.state_dict() for s in self.schedulers],
                    'optimizers': [o.state_dict() for o in self.optimizers]}
        torch.save(snapshot, os.path.join(self.opt['path']['training_state'], '{}.state'.format(iter_step)))

    def resume_training(self, resume_state, opt=None):
        for kind, mine in (('optimizers', self.optimizers), ('schedulers', self.schedulers)):
            saved = resume_state[kind]
            assert len(saved) == len(mine), 'Wrong lengths of {}'.format(kind)
            for obj, st in zip(mine, saved):
                obj.load_state_dict(st)
