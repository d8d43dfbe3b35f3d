import sys, torch
sys.path.insert(0, '.')
from oracle import dsn_oracle as D, srn_oracle as O
from dasr_b200.dsn.model import Discriminator
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
net = Discriminator(kernel_size=5, D_arch='FSD', norm_layer='Instance', filter_type='gau', cs='cat')
sd = O.synth_state_dict(D.fsd_shapes(3), 51, 1.0)
net.load_state_dict(sd, strict=False); net.cuda()
sdc = sd
def rel(a, b): return float((a.detach().cpu() - b.detach().cpu()).abs().max() / b.abs().max())
# 1) filter only
x = O.synth_image((2, 3, 24, 16), 52).cuda().requires_grad_(True)
y = net.filter(x); pat = O.synth(tuple(y.shape), 7).cuda(); (y * pat).sum().backward(); g1 = x.grad.clone()
x2 = x.detach().cpu().clone().requires_grad_(True); pat = pat.cpu()
y2 = D.fsd_filter(x2, 'gau'); (y2 * pat).sum().backward()
print('filter fwd', rel(y, y2), 'bwd', rel(g1, x2.grad))
# 2) net only
x = O.synth_image((2, 3, 24, 16), 53).cuda().requires_grad_(True)
y = net.net(x); pat = O.synth(tuple(y.shape), 8).cuda(); (y * pat).sum().backward(); g1 = x.grad.clone()
x2 = x.detach().cpu().clone().requires_grad_(True); pat = pat.cpu()
y2 = D.fsd_net(x2, sdc); (y2 * pat).sum().backward()
print('net fwd', rel(y, y2), 'bwd', rel(g1, x2.grad))
d = (g1.cpu() - x2.grad).abs()
print('max err at', torch.nonzero(d == d.max())[:3].tolist(), 'shape', tuple(d.shape))
print((d / x2.grad.abs().max())[0, 0, :4, :6])
