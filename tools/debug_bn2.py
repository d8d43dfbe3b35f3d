import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['DASR_B200_ALLOW_RANDOM_VGG'] = '1'
import torch, torch.nn as nn
from dasr_b200.dsn.model import DiscriminatorBasic
torch.manual_seed(0)
ours = DiscriminatorBasic(3, 'Batch').cuda().train()
ref = nn.Sequential(nn.Conv2d(3, 64, 5, padding=2), nn.LeakyReLU(0.2), nn.Conv2d(64, 128, 5, padding=2), nn.BatchNorm2d(128), nn.LeakyReLU(0.2),
                    nn.Conv2d(128, 256, 5, padding=2), nn.BatchNorm2d(256), nn.LeakyReLU(0.2), nn.Conv2d(256, 1, 1)).cuda().train()
with torch.no_grad():
    for k, v in ref.state_dict().items():
        if v.dtype.is_floating_point and 'running' not in k:
            v.copy_(torch.randn_like(v) * (0.05 if v.dim() > 1 else 0.3) + (1.0 if ('3.weight' in k or '6.weight' in k) else 0.0))
ours.net.load_state_dict(ref.state_dict())
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
x = torch.rand(3, 3, 20, 12, device='cuda')
xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
ya, yb = ours(xa), ref(xb)
pat = torch.randn_like(yb)
(ya * pat).sum().backward(); (yb * pat).sum().backward()
rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))
print('out', rel(ya, yb), 'dx', rel(xa.grad, xb.grad))
for (k, p), (_, q) in zip(ours.net.named_parameters(), ref.named_parameters()):
    print(k, tuple(p.shape), rel(p.grad, q.grad), float(q.grad.abs().max()))
