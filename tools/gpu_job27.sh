#!/bin/bash
cd "$(dirname "$0")/.."
for cfg in "DASR_B200_BATCH_SPLIT=1" "DASR_B200_BATCH_SPLIT=2" "DASR_B200_BATCH_SPLIT=4" "DASR_B200_BATCH_SPLIT=8" "DASR_B200_BATCH_SPLIT=16" "DASR_B200_BATCH_SPLIT=1"; do
  echo "== bench: $cfg"; env $cfg DASR_BENCH_FP16=0 timeout 900 python bench.py --train-steps 0 --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], 'frac', d['roofline']['frac'], d['clocks']['sm_mhz'], d['gpu_launches'])"
done
echo "== parity with split 2"; DASR_B200_BATCH_SPLIT=2 timeout 900 python -m pytest tests/test_gpu_parity_scale.py -q -s -m gpu -k config1 2>&1 | grep "config1 bf16 \|passed\|failed" | cut -c1-160
