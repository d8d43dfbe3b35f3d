#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== debug bn2"; timeout 120 python tools/debug_bn2.py 2>&1 | tail -14
echo "== f1 + dsn tests"; timeout 1500 python -m pytest tests/test_gpu_f1.py tests/test_gpu_dsn.py -q -s -m gpu > $O/r2_f1_tests.log 2>&1; grep -n "out .* dx\|passed\|failed\|FAILED\|Error\|assert " $O/r2_f1_tests.log | head -30
