"""DSN configs[4] iteration: standalone timing + per-kernel device time (torch profiler)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench

class A: train_steps = 3
torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
prec = os.environ.get('TRAIN_PREC', 'fp32')
res = bench.bench_dsn(A, dev, 0, 1, torch.cuda.synchronize, lambda ms: ms, prec)
print('standalone ms/iteration', res['ms_per_step'], flush=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    res = bench.bench_dsn(A, dev, 0, 1, torch.cuda.synchronize, lambda ms: ms, prec)
print('under profiler ms/iteration', res['ms_per_step'])
tot = {}
for e in prof.key_averages():
    tot[e.key] = (e.device_time_total if hasattr(e, 'device_time_total') else e.cuda_time_total, e.count)
s = sum(v[0] for v in tot.values())
for k, (t, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:14]:
    print('%-70s %9.1f ms %6d calls %5.1f%%' % (k[:70], t / 1e3, c, 100 * t / s))
print('total device ms (5 iterations):', s / 1e3)
