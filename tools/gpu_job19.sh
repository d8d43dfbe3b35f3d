#!/bin/bash
# final verification + evidence for the round: smoke, all GPU tests, bench (both arms), tail-layer ncu capture
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== gpu tests"; timeout 1800 python -m pytest tests -q -s -m gpu > $O/r2_gpu_tests.log 2>&1; grep -n "passed\|failed\|FAILED\|Error" $O/r2_gpu_tests.log | cut -c1-250 | head -12
echo "== bench (driver arguments)"; timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; cat $O/r2_bench_n1.json | cut -c1-3500
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2>/dev/null | cut -c1-900
echo "== ncu set full: upconvs + last layer"; NB=1 BATCH=16 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'conv_tc_kernel' -c 6 -o $O/r2_conv_tc_tail python tools/profile_forward.py > $O/r2_ncu_full_tail.log 2>&1; tail -2 $O/r2_ncu_full_tail.log
