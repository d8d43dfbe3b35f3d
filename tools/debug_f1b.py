"""Bisect the 1e-3 gradient deviation of Discriminator_VGG_128: run sub-chains conv{k}..conv4_1 (+BN+LeakyReLU) on random
inputs through (a) our layer-list kernels, (b) torch GPU fp32, (c) float64 CPU, and print dx / dw errors per chain."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn
from dasr_b200 import seqnet

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)


class Chain(nn.Module):
    def __init__(self, specs):
        super().__init__()
        self.names = []
        for i, (ci, co, k, s, bn) in enumerate(specs):
            setattr(self, 'conv%d' % i, nn.Conv2d(ci, co, k, s, 1, bias=not bn))
            if bn:
                setattr(self, 'bn%d' % i, nn.BatchNorm2d(co))
            self.names.append((i, bn))

    def seq(self):
        act = nn.LeakyReLU(0.2)
        out = []
        for i, bn in self.names:
            out.append(('conv%d' % i, getattr(self, 'conv%d' % i)))
            if bn:
                out.append(('bn%d' % i, getattr(self, 'bn%d' % i)))
            out.append(('_a%d' % i, act))
        return out

    def forward(self, x, native=False):
        if native:
            for _, m in self.seq():
                x = m(x)
            return x
        return seqnet.run_module(self, x, seqnet.compile_sequence(self.seq(), '', self.training))


FULL = [(3, 64, 3, 1, False), (64, 64, 4, 2, True), (64, 128, 3, 1, True), (128, 128, 4, 2, True), (128, 256, 3, 1, True),
        (256, 256, 4, 2, True), (256, 512, 3, 1, True), (512, 512, 4, 2, True), (512, 512, 3, 1, True), (512, 512, 4, 2, True)]
SIZES = [128, 128, 64, 64, 32, 32, 16, 16, 8, 8]
rl = lambda a, b: float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp_min(1e-30))
cases = [('chain from layer %d' % k, FULL[k:], SIZES[k]) for k in (9, 8, 7, 6, 5, 4, 2, 0)]
cases += [('single layer %d' % k, FULL[k:k + 1], SIZES[k]) for k in (1, 2, 3, 4, 6)]
for title, specs, size in cases:
    net = Chain(specs).train()
    with torch.no_grad():
        for k, p in net.named_parameters():
            if 'bn' in k and k.endswith('weight'):
                p.copy_(1.0 + 0.3 * torch.randn_like(p))
            elif 'bn' in k:
                p.copy_(0.05 * torch.randn_like(p))
    x = torch.rand(2, specs[0][0], size, size)
    n64 = copy.deepcopy(net).double()
    x64 = x.double().requires_grad_(True)
    o64 = n64(x64, native=True)
    pat = torch.randn_like(o64)
    (o64 * pat).sum().backward()
    res = {}
    for tag, native in (('ours', False), ('torch', True)):
        m = copy.deepcopy(net).cuda()
        xx = x.cuda().requires_grad_(True)
        o = m(xx, native=native)
        (o * pat.float().cuda()).sum().backward()
        w0 = dict(m.named_parameters())['conv0.weight'].grad
        res[tag] = (rl(o, o64), rl(xx.grad, x64.grad), rl(w0, dict(n64.named_parameters())['conv0.weight'].grad))
    print('%-22s out %.1e / %.1e   dx %.1e / %.1e   dW(first) %.1e / %.1e     (ours / torch-GPU vs float64)'
          % (title, res['ours'][0], res['torch'][0], res['ours'][1], res['torch'][1], res['ours'][2], res['torch'][2]))
