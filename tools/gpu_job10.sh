#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== all gpu tests"; timeout 1800 python -m pytest tests -q -s -m gpu > $O/r2_gpu_tests.log 2>&1; grep -n "passed\|failed\|FAILED\|Error" $O/r2_gpu_tests.log | head
echo "== BN discriminator fwd+bwd timing (16x3x128x128)"; timeout 300 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from dasr_b200.srn.models.modules.architecture import Discriminator_VGG_128
d = Discriminator_VGG_128(3, 64).cuda().train()
x = torch.rand(16, 3, 128, 128, device='cuda', requires_grad=True)
for _ in range(3):
    d(x).sum().backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    d(x).sum().backward()
e1.record(); torch.cuda.synchronize()
print('Discriminator_VGG_128 fwd+bwd: %.2f ms' % (e0.elapsed_time(e1) / 10))
PY
echo "== bench"; timeout 900 python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; cat $O/r2_bench_n1.json | cut -c1-4000
