#!/bin/bash
# round-2 GPU job: pair-kernel bring-up + probes + parity at scale + timings
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=dasr_b200/lib
echo "== selftest check"; timeout 600 $L/selftest check > $O/r2_selftest_check.log 2>&1; RC=$?; echo "rc=$RC"; grep -c "PASS" $O/r2_selftest_check.log; grep "FAIL" $O/r2_selftest_check.log | head -10; tail -2 $O/r2_selftest_check.log
if [ $RC -ne 0 ]; then export DASR_B200_PAIR=0; echo "PAIR KERNEL DISABLED for the rest of the job"; fi
echo "== fused"; timeout 300 $L/selftest fused 2>&1 | tee $O/r2_fused.log | tail -10
echo "== stage times pair=$DASR_B200_PAIR"; timeout 300 python tools/stage_times.py 2>&1 | tail -2 | tee $O/r2_stage_times_pair1.log
echo "== stage times pdl=0"; DASR_B200_PDL=0 timeout 300 python tools/stage_times.py 2>&1 | tail -2 | tee $O/r2_stage_times_pdl0.log
echo "== scale tests"; timeout 900 python -m pytest tests/test_gpu_parity_scale.py -q -s -m gpu > $O/r2_scale_tests.log 2>&1; grep "config\|mixed\|passed\|failed\|Error" $O/r2_scale_tests.log | head -20
echo "== gpu tests"; timeout 1500 python -m pytest tests -q -m gpu > $O/r2_gpu_tests.log 2>&1; tail -8 $O/r2_gpu_tests.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2_bench1.log 2>&1; tail -c 700 $O/r2_bench1.log
