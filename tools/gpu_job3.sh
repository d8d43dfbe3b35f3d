#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=dasr_b200/lib
echo "== selftest check"; timeout 600 $L/selftest check > $O/r2_selftest_check.log 2>&1; RC=$?; echo "rc=$RC"; grep -c "PASS" $O/r2_selftest_check.log; grep "FAIL" $O/r2_selftest_check.log | head -10
echo "== debug bn"; timeout 120 python tools/debug_bn.py 2>&1 | tail -3
echo "== gpu tests"; timeout 1500 python -m pytest tests -q -m gpu > $O/r2_gpu_tests.log 2>&1; tail -6 $O/r2_gpu_tests.log
echo "== bench full"; timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench1.log 2>$O/r2_bench1.err; cat $O/r2_bench1.log | tail -c 2500
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
