#!/bin/bash
# round-2 GPU job: pair kernel with tail blocks, all dense-block launches on pairs
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=dasr_b200/lib
echo "== selftest check"; timeout 600 $L/selftest check > $O/r2_selftest_check.log 2>&1; RC=$?; echo "rc=$RC"; grep -c "PASS" $O/r2_selftest_check.log; grep "FAIL" $O/r2_selftest_check.log | head -10; tail -2 $O/r2_selftest_check.log
if [ $RC -ne 0 ]; then export DASR_B200_PAIR=0; echo "PAIR KERNEL DISABLED for the rest of the job"; fi
echo "== fused"; timeout 300 $L/selftest fused 2>&1 | tee $O/r2_fused.log | grep pair
echo "== fused nblk=3"; DASR_TC2_NBLK=3 timeout 300 $L/selftest fused 2>&1 | grep pair | tee $O/r2_fused_nblk3.log
echo "== stage times"; timeout 300 python tools/stage_times.py 2>&1 | tail -2 | tee $O/r2_stage_times.log
echo "== gpu tests"; timeout 1500 python -m pytest tests -q -m gpu > $O/r2_gpu_tests.log 2>&1; tail -6 $O/r2_gpu_tests.log
echo "== bench sched3"; timeout 600 python bench.py --steps 10 --warmup 3 --train-steps 0 --no-cpu-baseline > $O/r2_bench_s3.log 2>$O/r2_bench_s3.err; tail -c 600 $O/r2_bench_s3.log
echo "== bench sched2"; DASR_B200_SCHED=2 timeout 600 python bench.py --steps 10 --warmup 3 --train-steps 0 --no-cpu-baseline > $O/r2_bench_s2.log 2>/dev/null; tail -c 300 $O/r2_bench_s2.log
echo "== train phases"; timeout 600 python tools/time_train_phases.py > $O/r2_train_phases.log 2>&1; tail -25 $O/r2_train_phases.log
echo "== bench full"; timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2_bench1.log 2>$O/r2_bench1.err; cat $O/r2_bench1.log | tail -c 1500
