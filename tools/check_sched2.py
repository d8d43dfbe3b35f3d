"""Dense-block schedule 2 (DASR_B200_SCHED=2) against schedule 1 and the fp32 oracle: error + timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import srn_oracle as O
from dasr_b200 import engine

def run(nb, shape, sched):
    os.environ['DASR_B200_SCHED'] = sched
    sd = O.synth_state_dict(O.rrdbnet_shapes(nb=nb), 1, 0.1)
    params = [v.cuda() for v in sd.values()]
    x = O.synth_image(shape, 2).cuda()
    cache = engine._PackCache()
    out = engine.rrdb_forward_bf16(x, params, nb, 4, cache)
    torch.cuda.synchronize()
    return out, params, x, cache, sd

nb, shape = 1, (2, 3, 40, 24)
a, _, x, _, sd = run(nb, shape, '1')
b, *_ = run(nb, shape, '2')
ref = O.rrdbnet_forward(x.cpu(), sd, nb)
rel = lambda u: float((u.float().cpu() - ref).abs().max() / ref.abs().max())
print('rel-Linf vs fp32 oracle: sched1 %.3e  sched2 %.3e' % (rel(a), rel(b)), flush=True)
nb, shape = 23, (16, 3, 256, 256)
for sched in ('1', '2'):
    out, params, x, cache, _ = run(nb, shape, sched)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        engine.rrdb_forward_bf16(x, params, nb, 4, cache)
    e0.record()
    for _ in range(5):
        engine.rrdb_forward_bf16(x, params, nb, 4, cache)
    e1.record()
    torch.cuda.synchronize()
    print('schedule', sched, 'ms/forward %.2f' % (e0.elapsed_time(e1) / 5), flush=True)
