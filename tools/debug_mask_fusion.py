"""Gradients of the mixed-precision RRDBNet backward with the LeakyReLU masks applied by act_bwd kernels (FUSE_MASK=0) vs inside
the dgrad epilogues (FUSE_MASK=1), and both against the fp32 backward: per-parameter rel-L2."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dasr_b200 import engine
from oracle import srn_oracle as O

nb = 2
sd = O.synth_state_dict(O.rrdbnet_shapes(nb=nb), 5, 0.3)
params = [v.cuda() for v in sd.values()]
x = O.synth_image((4, 3, 32, 32), 6).cuda()
dout = O.synth((4, 3, 128, 128), 7).cuda()
res = {}
for mode in ('0', '1'):
    os.environ['DASR_B200_FUSE_MASK'] = mode
    out, ctx = engine.rrdb_forward_bf16_train(x, params, nb, 4, engine._PackCache())
    _, grads, _ = engine.rrdb_backward_bf16(ctx, params, dout, engine._PackCache())
    torch.cuda.synchronize()
    res[mode] = [g.clone() for g in grads]
o32, c32 = engine.rrdb_forward_f32(x, params, nb, 4, save=True)
g32 = engine.rrdb_backward_f32(c32, params, dout)[1] if hasattr(engine, 'rrdb_backward_f32') else None
names = list(sd.keys())
worst = 0.0
for i, n in enumerate(names):
    a, b = res['0'][i].double(), res['1'][i].double()
    d = float((a - b).norm() / a.norm().clamp_min(1e-30))
    line = '%-40s fused-vs-unfused %.3e' % (n, d)
    if g32 is not None and g32[i] is not None:
        r = g32[i].double()
        line += '   unfused-vs-fp32 %.3e   fused-vs-fp32 %.3e' % (float((a - r).norm() / r.norm().clamp_min(1e-30)), float((b - r).norm() / r.norm().clamp_min(1e-30)))
    worst = max(worst, d)
    if i < 14 or d > 5e-2:
        print(line)
print('worst fused-vs-unfused rel-L2: %.3e' % worst)
