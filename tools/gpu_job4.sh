#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== gpu tests"; timeout 1800 python -m pytest tests -q -s -m gpu > $O/r2_gpu_tests.log 2>&1; grep -n "out .* dx\|passed\|failed\|FAILED\|Error\|dp2\|config\|mixed" $O/r2_gpu_tests.log | head -40
