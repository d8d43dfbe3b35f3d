#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
echo "== selftest"; timeout 600 dasr_b200/lib/selftest check > $O/r2_selftest_check.log 2>&1; grep -c PASS $O/r2_selftest_check.log; grep "FAIL" $O/r2_selftest_check.log | head -30; grep "epi7" $O/r2_selftest_check.log | head -5
echo "== mixed-precision tests"; timeout 900 python -m pytest tests -q -s -m gpu -k "mixed or bf16 or train_steps or packer or dp" 2>&1 | grep -n "passed\|failed\|FAILED\|mixed-prec\|Error" | cut -c1-300 | head
for cfg in "DASR_B200_FUSE_MASK=0" "DASR_B200_FUSE_MASK=1" "DASR_B200_FUSE_MASK=0" "DASR_B200_FUSE_MASK=1"; do
  echo "== train: $cfg"; env $cfg TRAIN_PREC=bf16 STEPS=20 timeout 600 python tools/one_train_step.py 2>&1 | tail -1
done
