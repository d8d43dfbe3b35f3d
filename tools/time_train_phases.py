"""Per-kernel-family device time of one DASR train step (torch profiler, CUDA activities only)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench

class A: train_steps = 1
prec = os.environ.get('TRAIN_PREC', 'bf16')
torch.cuda.set_device(0)
dev = torch.device('cuda', 0)

def barrier(): torch.cuda.synchronize()

# reuse bench_train's model construction by monkeypatching its loop: run it once to warm up
res = bench.bench_train(A, dev, 0, 1, barrier, lambda ms: ms, prec)
print('warm step ms', res['ms_per_step'])
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    res = bench.bench_train(A, dev, 0, 1, barrier, lambda ms: ms, prec)
tot = {}
for e in prof.key_averages():
    tot[e.key] = (e.device_time_total if hasattr(e, 'device_time_total') else e.cuda_time_total, e.count)
s = sum(v[0] for v in tot.values())
for k, (t, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:18]:
    print('%-70s %9.1f ms %6d calls %5.1f%%' % (k[:70], t / 1e3, c, 100 * t / s))
print('total device ms (3 steps incl. warmup inside bench_train):', s / 1e3)
