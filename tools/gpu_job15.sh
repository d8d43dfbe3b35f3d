#!/bin/bash
# refresh the forward evidence for the final build: ncu launch list (+DRAM bytes) of one forward, traffic json, racecheck
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=dasr_b200/lib
echo "== ncu launch list"; timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2_launches_forward_bf16_final.csv python tools/profile_forward.py > $O/r2_ncu_list.log 2>&1; tail -2 $O/r2_ncu_list.log
python tools/summarize_launches.py $O/r2_launches_forward_bf16_final.csv $O/r2_traffic.json 2>&1 | tail -14
echo "== ncu set full: last layer (taps in N) and second upconv"; NB=1 BATCH=16 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'conv_tc_kernel' -c 12 -o $O/r2_conv_tc_tail python tools/profile_forward.py > $O/r2_ncu_full_tail.log 2>&1; tail -3 $O/r2_ncu_full_tail.log
ls -la $O/r2_conv_tc_tail.ncu-rep
echo "== racecheck (small shapes)"; timeout 1200 compute-sanitizer --tool racecheck $L/selftest check > $O/r2_racecheck_selftest.log 2>&1; tail -4 $O/r2_racecheck_selftest.log
