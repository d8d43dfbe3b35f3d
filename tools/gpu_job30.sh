#!/bin/bash
cd "$(dirname "$0")/.."
for cfg in "DASR_F32_THIN=0" "DASR_F32_THIN=1" "DASR_F32_THIN=0" "DASR_F32_THIN=1"; do
echo "== dsn $cfg"; env $cfg timeout 600 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
os.environ.setdefault('DASR_B200_ALLOW_RANDOM_VGG', '1')
import bench
class A: train_steps = 20
torch.cuda.set_device(0)
r = bench.bench_dsn(A, torch.device('cuda', 0), 0, 1, torch.cuda.synchronize, lambda ms: ms, 'bf16')
print('dsn bf16 ms', r['ms_per_step'])
PY
echo "== train $cfg"; env $cfg TRAIN_PREC=bf16 STEPS=20 timeout 600 python tools/one_train_step.py 2>&1 | tail -1
done
