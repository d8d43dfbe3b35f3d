#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== selftest"; timeout 600 dasr_b200/lib/selftest check > $O/r2_selftest_check.log 2>&1; grep -c PASS $O/r2_selftest_check.log; grep "FAIL" $O/r2_selftest_check.log | head -20; grep "tf32" $O/r2_selftest_check.log | head -8
for cfg in "DASR_B200_BWD_OVERLAP=0 DASR_B200_SIDE_MATH=fma" "DASR_B200_BWD_OVERLAP=1 DASR_B200_SIDE_MATH=fma" "DASR_B200_BWD_OVERLAP=1 DASR_B200_SIDE_MATH=tf32" "DASR_B200_BWD_OVERLAP=1 DASR_B200_SIDE_MATH=tf32 DASR_B200_PAIR_DGRAD=1"; do
  echo "== mixed train step: $cfg"; env $cfg TRAIN_PREC=bf16 STEPS=20 timeout 600 python tools/one_train_step.py 2>&1 | tail -1
done
for cfg in "DASR_B200_F32_MATH=fma" "DASR_B200_F32_MATH=tf32x3" "DASR_B200_F32_MATH=tf32"; do
  echo "== fp32-mode train step: $cfg"; env $cfg TRAIN_PREC=fp32 STEPS=3 timeout 600 python tools/one_train_step.py 2>&1 | tail -1
done
echo "== all gpu tests (default math)"; timeout 1800 python -m pytest tests -q -s -m gpu > $O/r2_gpu_tests.log 2>&1; grep -n "passed\|failed\|FAILED\|mixed-prec\|Error" $O/r2_gpu_tests.log | head
echo "== gpu tests with the fp32 kernels on 3xTF32"; DASR_B200_F32_MATH=tf32x3 timeout 1800 python -m pytest tests -q -s -m gpu > $O/r2_gpu_tests_tf32x3.log 2>&1; grep -n "passed\|failed\|FAILED\|config0\|Error" $O/r2_gpu_tests_tf32x3.log | head -30
echo "== pair dgrad parity"; DASR_B200_PAIR_DGRAD=1 timeout 900 python -m pytest tests -q -s -m gpu -k "mixed or bf16" 2>&1 | grep -n "passed\|failed\|FAILED\|mixed-prec\|Error" | head
echo "== tiny torch ops"; timeout 600 python tools/find_small_ops.py 2>&1 | tail -45
