#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
echo "== gpu tests"; timeout 1800 python -m pytest tests -q -s -m gpu > $O/r2_gpu_tests.log 2>&1; grep -n "passed\|failed\|FAILED\|Error" $O/r2_gpu_tests.log | cut -c1-250 | head -12
