#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== selftest"; timeout 600 dasr_b200/lib/selftest check > $O/r2_selftest_check.log 2>&1; grep -c PASS $O/r2_selftest_check.log; grep "FAIL" $O/r2_selftest_check.log | head -20; grep "kind2.*mode0" $O/r2_selftest_check.log | head
echo "== gpu tests"; timeout 1800 python -m pytest tests -q -s -m gpu > $O/r2_gpu_tests.log 2>&1; grep -n "passed\|failed\|FAILED\|Error" $O/r2_gpu_tests.log | head
for cfg in "DASR_B200_UP_STAGED=0" "DASR_B200_UP_STAGED=1"; do
  echo "== bench: $cfg"; env $cfg timeout 900 python bench.py --train-steps 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
echo "== dsn"; timeout 600 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
os.environ.setdefault('DASR_B200_ALLOW_RANDOM_VGG', '1')
import bench
class A: train_steps = 20
torch.cuda.set_device(0)
r = bench.bench_dsn(A, torch.device('cuda', 0), 0, 1, torch.cuda.synchronize, lambda ms: ms, 'bf16')
print('dsn ms', r['ms_per_step'])
PY
echo "== train"; TRAIN_PREC=bf16 STEPS=20 timeout 600 python tools/one_train_step.py 2>&1 | tail -1
