import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasr_b200 import ops
torch.manual_seed(0)
N, H, W, C = 3, 20, 12, 128
x = torch.randn(N, H, W, C, device='cuda')
gamma = torch.rand(C, device='cuda') + 0.5
beta = torch.randn(C, device='cuda') * 0.1
rm, rv = torch.zeros(C, device='cuda'), torch.ones(C, device='cuda')
st = torch.empty(C, 2, device='cuda')
y = torch.empty_like(x)
ops.bn_lrelu_fwd(x, y, gamma, beta, rm, rv, st, 1e-5, 0.1, True, 0.2)
xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
bn = torch.nn.BatchNorm2d(C).cuda()
bn.weight.data.copy_(gamma); bn.bias.data.copy_(beta)
yr = torch.nn.functional.leaky_relu(bn(xr), 0.2)
print('fwd', float((y.permute(0, 3, 1, 2) - yr).abs().max()), 'rm', float((rm - bn.running_mean).abs().max()), 'rv', float((rv - bn.running_var).abs().max()))
dy = torch.randn_like(y)
(yr * dy.permute(0, 3, 1, 2)).sum().backward()
dx = torch.empty_like(x); dg = torch.empty(C, device='cuda'); db = torch.empty(C, device='cuda')
ops.bn_lrelu_bwd(x, y, dy, gamma, st, dx, dg, db, True, 0.2)
print('dx', float((dx.permute(0, 3, 1, 2) - xr.grad).abs().max() / xr.grad.abs().max()), 'dgamma', float((dg - bn.weight.grad).abs().max() / bn.weight.grad.abs().max()),
      'dbeta', float((db - bn.bias.grad).abs().max() / bn.bias.grad.abs().max()))
