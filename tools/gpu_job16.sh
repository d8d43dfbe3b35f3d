#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== selftest"; timeout 600 dasr_b200/lib/selftest check > $O/r2_selftest_check.log 2>&1; grep -c PASS $O/r2_selftest_check.log; grep "FAIL" $O/r2_selftest_check.log | head -30
echo "== selftest without chunk64"; DASR_TC_CHUNK64=0 timeout 600 dasr_b200/lib/selftest check 2>&1 | grep -c PASS
echo "== gpu tests"; timeout 1800 python -m pytest tests -q -s -m gpu > $O/r2_gpu_tests.log 2>&1; grep -n "passed\|failed\|FAILED\|Error" $O/r2_gpu_tests.log | cut -c1-250 | head -12
for cfg in "DASR_TC_CHUNK64=0" "DASR_TC_CHUNK64=1"; do
  echo "== bench: $cfg"; env $cfg timeout 900 python bench.py --train-steps 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks']['sm_mhz'], 'fp16', d.get('fp16',{}).get('ms_per_step'))"
  echo "== train: $cfg"; env $cfg TRAIN_PREC=bf16 STEPS=20 timeout 600 python tools/one_train_step.py 2>&1 | tail -1
done
