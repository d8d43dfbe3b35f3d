#!/bin/bash
# round-2 GPU job 1: pair-kernel bring-up + probes + parity at scale + baseline timings
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=dasr_b200/lib
echo "== mma2rate"; timeout 120 $L/selftest mma2rate 2>&1 | tee $O/r2_mma2rate.log
echo "== selftest check"; timeout 600 $L/selftest check > $O/r2_selftest_check.log 2>&1; echo "rc=$?"; grep -c "\[ok\]\|\[OK\]" $O/r2_selftest_check.log; grep -i "fail\|error" $O/r2_selftest_check.log | head -20; tail -3 $O/r2_selftest_check.log
echo "== fused"; timeout 300 $L/selftest fused 2>&1 | tee $O/r2_fused.log
echo "== tmarate"; timeout 300 $L/selftest tmarate 2>&1 | tee $O/r2_tmarate.log
echo "== stage times pair=1"; DASR_B200_PAIR=1 timeout 300 python tools/stage_times.py 2>&1 | tail -3 | tee $O/r2_stage_times_pair1.log
echo "== stage times pair=0"; DASR_B200_PAIR=0 timeout 300 python tools/stage_times.py 2>&1 | tail -3 | tee $O/r2_stage_times_pair0.log
echo "== stage times pair=1 pdl=0"; DASR_B200_PDL=0 timeout 300 python tools/stage_times.py 2>&1 | tail -3 | tee $O/r2_stage_times_pdl0.log
echo "== scale tests"; timeout 900 python -m pytest tests/test_gpu_parity_scale.py -x -q -s -m gpu > $O/r2_scale_tests.log 2>&1; grep "config\|mixed\|passed\|failed\|Error" $O/r2_scale_tests.log | head -20
echo "== gpu tests"; timeout 1200 python -m pytest tests -x -q -m gpu > $O/r2_gpu_tests.log 2>&1; tail -5 $O/r2_gpu_tests.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2_bench1.log 2>&1; tail -c 600 $O/r2_bench1.log
