#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== selftest"; timeout 600 dasr_b200/lib/selftest check > $O/r2_selftest_check.log 2>&1; grep -c PASS $O/r2_selftest_check.log; grep "FAIL" $O/r2_selftest_check.log | head -30
echo "== gpu tests"; timeout 1800 python -m pytest tests -q -s -m gpu > $O/r2_gpu_tests.log 2>&1; grep -n "passed\|failed\|FAILED\|Error" $O/r2_gpu_tests.log | cut -c1-250 | head -12
echo "== train mixed"; TRAIN_PREC=bf16 STEPS=20 timeout 600 python tools/one_train_step.py 2>&1 | tail -1
echo "== train fp32"; TRAIN_PREC=fp32 STEPS=3 timeout 600 python tools/one_train_step.py 2>&1 | tail -1
echo "== dsn"; timeout 600 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
os.environ.setdefault('DASR_B200_ALLOW_RANDOM_VGG', '1')
import bench
class A: train_steps = 20
torch.cuda.set_device(0)
r = bench.bench_dsn(A, torch.device('cuda', 0), 0, 1, torch.cuda.synchronize, lambda ms: ms, 'bf16')
print('dsn bf16 ms', r['ms_per_step'])
PY
