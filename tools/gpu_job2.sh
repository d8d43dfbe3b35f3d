#!/bin/bash
# round-2 GPU job: verify TMA-warp fix, tests, bench, ncu evidence (launch list with DRAM bytes + set-full of one dense block)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=dasr_b200/lib
echo "== selftest check"; timeout 600 $L/selftest check > $O/r2_selftest_check.log 2>&1; RC=$?; echo "rc=$RC"; grep -c "PASS" $O/r2_selftest_check.log; grep "FAIL" $O/r2_selftest_check.log | head -10
if [ $RC -ne 0 ]; then export DASR_B200_PAIR=0; echo "PAIR KERNEL DISABLED for the rest of the job"; fi
echo "== fused"; timeout 300 $L/selftest fused 2>&1 | tee $O/r2_fused.log | grep pair
echo "== stage times"; timeout 300 python tools/stage_times.py 2>&1 | tail -2 | tee $O/r2_stage_times.log
echo "== gpu tests"; timeout 1500 python -m pytest tests -q -m gpu > $O/r2_gpu_tests.log 2>&1; tail -6 $O/r2_gpu_tests.log
echo "== bench sched3"; timeout 600 python bench.py --steps 20 --warmup 5 --train-steps 0 --no-cpu-baseline > $O/r2_bench_s3.log 2>$O/r2_bench_s3.err; tail -c 700 $O/r2_bench_s3.log
echo "== ncu launch list"; timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2_launches_forward_bf16_sched3.csv python tools/profile_forward.py > $O/r2_ncu_list.log 2>&1; tail -2 $O/r2_ncu_list.log
python tools/summarize_launches.py $O/r2_launches_forward_bf16_sched3.csv $O/r2_traffic.json 2>&1 | tail -12
echo "== ncu set full (dense block 10: launches 51..55 of the conv kernels)"; NB=4 BATCH=16 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc -s 11 -c 5 -o $O/r2_conv_tc_sched3_rdb3 python tools/profile_forward.py > $O/r2_ncu_full.log 2>&1; tail -3 $O/r2_ncu_full.log
echo "== racecheck (small shapes)"; timeout 900 compute-sanitizer --tool racecheck $L/selftest check > $O/r2_racecheck_selftest.log 2>&1; tail -4 $O/r2_racecheck_selftest.log
