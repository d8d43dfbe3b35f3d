#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
echo "== selftest"; timeout 600 dasr_b200/lib/selftest check > $O/r2_selftest_check.log 2>&1; grep -c PASS $O/r2_selftest_check.log; grep "FAIL" $O/r2_selftest_check.log | head -30
echo "== scale tests"; timeout 900 python -m pytest tests/test_gpu_parity_scale.py tests/test_gpu_parity.py -q -s -m gpu 2>&1 | grep -n "passed\|failed\|FAILED\|config1" | cut -c1-200 | head
for cfg in "DASR_B200_TILE_REV=0" "DASR_B200_TILE_REV=1" "DASR_B200_TILE_REV=0" "DASR_B200_TILE_REV=1"; do
  echo "== bench: $cfg"; env $cfg DASR_BENCH_FP16=0 timeout 900 python bench.py --train-steps 0 --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], 'frac', d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
