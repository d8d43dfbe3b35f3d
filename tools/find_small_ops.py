"""Which Python lines launch the many tiny torch kernels (fills, copies) of a mixed-precision DASR train step?
torch profiler with stacks over ONE eager step (graphs off), grouped by the innermost dasr_b200 / bench frame."""
import os
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault('DASR_B200_ALLOW_RANDOM_VGG', '1')
import torch
from torch.profiler import profile, ProfilerActivity
import bench

GRAPH = os.environ.get('DASR_B200_GRAPH', '1')


class A:
    train_steps = 1


dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
holder = {}
orig = bench.time_steps if hasattr(bench, 'time_steps') else None
prec = os.environ.get('TRAIN_PREC', 'bf16')
# build the model through bench_train once (warm), then profile a second call's timed step only via the profiler schedule
bench.bench_train(A, dev, 0, 1, torch.cuda.synchronize, lambda ms: ms, prec)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    bench.bench_train(A, dev, 0, 1, torch.cuda.synchronize, lambda ms: ms, prec)
want = ('aten::zero_', 'aten::fill_', 'aten::zeros', 'aten::copy_', 'aten::zeros_like', 'aten::clone', 'aten::cat', 'aten::mul', 'aten::add')
cnt = Counter()
for e in prof.events():
    if e.name in want and e.stack:
        fr = [f for f in e.stack if ('dasr_b200' in f or 'bench.py' in f)]
        cnt[(e.name, fr[0] if fr else e.stack[0])] += 1
for (name, frame), n in cnt.most_common(40):
    print('%6d  %-16s %s' % (n, name, frame[-110:]))
