"""Run the configs[2] train step a few times (for ncu / profiler captures): TRAIN_PREC=bf16|fp32, STEPS=n."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

class A: train_steps = int(os.environ.get('STEPS', 1))
torch.cuda.set_device(0)
res = bench.bench_train(A, torch.device('cuda', 0), 0, 1, torch.cuda.synchronize, lambda ms: ms, os.environ.get('TRAIN_PREC', 'bf16'))
print(res['ms_per_step'])
