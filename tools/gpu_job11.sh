#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== D tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dsn.py tests/test_gpu_parity_scale.py -q -s -m gpu -k "nlayer or dasr_model or fsd or discriminator or mixed" 2>&1 | grep -v Warning | grep -n "passed\|failed\|FAILED\|Error\|assert\|mixed-prec" | head -20
for cfg in "DASR_B200_FUSED_IN=0" "DASR_B200_FUSED_IN=1"; do
  echo "== mixed train step: $cfg"; env $cfg TRAIN_PREC=bf16 STEPS=20 timeout 600 python tools/one_train_step.py 2>&1 | tail -1
done
