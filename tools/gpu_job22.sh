#!/bin/bash
cd "$(dirname "$0")/.."
echo "== schedule tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -s -m gpu -k "schedules_agree" 2>&1 | tail -4
for cfg in "DASR_B200_SCHED=3" "DASR_B200_SCHED=4" "DASR_B200_SCHED=3" "DASR_B200_SCHED=4"; do
  echo "== bench: $cfg"; env $cfg DASR_BENCH_FP16=0 timeout 900 python bench.py --train-steps 0 --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], 'frac', d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
