#!/bin/bash
cd "$(dirname "$0")/.."
echo "== bench"; DASR_BENCH_FP16=0 timeout 900 python bench.py --train-steps 0 --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks']['sm_mhz'])"
echo "== api tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "sr_model or dasr_model or chop or x8" 2>&1 | tail -2
