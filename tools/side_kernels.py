"""Short workload for `ncu --set full` of the round-2 side kernels: the one-kernel discriminator layer (tf32), the tf32 generic conv
(forward / dgrad / wgrad) on discriminator shapes, the split BatchNorm kernels, the reworked bias-gradient reduction."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('DASR_B200_ALLOW_RANDOM_VGG', '1')
import torch
from dasr_b200 import ops
from dasr_b200.srn.models.modules.architecture import NLayerDiscriminator, Discriminator_VGG_128

torch.cuda.set_device(0)
d = NLayerDiscriminator(9, n_layers=3).cuda().train()
x = torch.rand(32, 9, 64, 64, device='cuda', requires_grad=True)
with ops.f32_math('tf32'):
    for _ in range(2):
        d(x).sum().backward()
v = Discriminator_VGG_128(3, 64).cuda().train()
y = torch.rand(16, 3, 128, 128, device='cuda', requires_grad=True)
for _ in range(2):
    v(y).sum().backward()
g = torch.randn(32, 32, 32, 192, device='cuda').to(torch.bfloat16)
db = torch.empty(128, device='cuda')
for _ in range(2):
    ops.bias_grad(ops.View(g, 128, 64), db)
torch.cuda.synchronize()
print('done')
