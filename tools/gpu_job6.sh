#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== selftest"; timeout 600 dasr_b200/lib/selftest check > $O/r2_selftest_check.log 2>&1; grep -c PASS $O/r2_selftest_check.log; grep "FAIL\|f16" $O/r2_selftest_check.log | head -20
echo "== gpu tests"; timeout 1800 python -m pytest tests -q -s -m gpu > $O/r2_gpu_tests.log 2>&1; grep -n "config\|mixed\|passed\|failed\|FAILED\|Error" $O/r2_gpu_tests.log | head -40
echo "== bench"; timeout 900 python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; cat $O/r2_bench_n1.json
