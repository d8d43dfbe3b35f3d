#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== selftest"; timeout 600 dasr_b200/lib/selftest check > $O/r2_selftest_check.log 2>&1; grep -c PASS $O/r2_selftest_check.log; grep "FAIL" $O/r2_selftest_check.log | head -20; grep "epi6" $O/r2_selftest_check.log | head
echo "== gpu tests"; timeout 1800 python -m pytest tests -q -s -m gpu > $O/r2_gpu_tests.log 2>&1; grep -n "passed\|failed\|FAILED\|Error\|config" $O/r2_gpu_tests.log | cut -c1-250 | head -12
for cfg in "DASR_B200_TAPN=0" "DASR_B200_TAPN=1"; do
  echo "== bench: $cfg"; env $cfg timeout 900 python bench.py --train-steps 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks']['sm_mhz'], 'fp16', d.get('fp16',{}).get('ms_per_step'))"
done
