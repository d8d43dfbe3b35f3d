"""Persistent-RDB kernel (DASR_B200_RDB=1) against the five-launch fused schedule: bitwise equality + timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import srn_oracle as O
from dasr_b200 import engine

def run(nb, shape, flag, chunk=2):
    os.environ['DASR_B200_RDB'] = flag
    engine.RDB_CHUNK_IMGS = chunk
    sd = O.synth_state_dict(O.rrdbnet_shapes(nb=nb), 1, 0.1)
    params = [v.cuda() for v in sd.values()]
    x = O.synth_image(shape, 2).cuda()
    cache = engine._PackCache()
    out = engine.rrdb_forward_bf16(x, params, nb, 4, cache)
    torch.cuda.synchronize()
    return out, params, x, cache

for nb, shape, chunk in ((1, (3, 3, 32, 32), 2), (1, (2, 3, 40, 24), 1), (2, (5, 3, 64, 64), 2), (1, (4, 3, 128, 128), 3)):
    a, *_ = run(nb, shape, '0')
    b, *_ = run(nb, shape, '1', chunk)
    print('nb', nb, shape, 'chunk', chunk, 'equal', bool(torch.equal(a, b)), 'max diff', float((a - b).abs().max()), flush=True)

if os.environ.get('TIME', '1') == '1':
    nb, shape = 23, (16, 3, 256, 256)
    for flag, chunk in (('0', 2), ('1', 2), ('1', 1), ('1', 4)):
        out, params, x, cache = run(nb, shape, flag, chunk)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            engine.rrdb_forward_bf16(x, params, nb, 4, cache)
        e0.record()
        for _ in range(3):
            engine.rrdb_forward_bf16(x, params, nb, 4, cache)
        e1.record()
        torch.cuda.synchronize()
        print('RDB kernel', flag, 'chunk', chunk, 'ms/forward %.2f' % (e0.elapsed_time(e1) / 3), flush=True)
        if flag == '0':
            ref = out
        else:
            print('   equal to five-launch schedule:', bool(torch.equal(ref, out)), flush=True)
