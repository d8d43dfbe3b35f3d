#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TRAIN_PREC=bf16 timeout 600 python tools/time_train_phases.py 2>&1 | grep -v Warn | tail -22 | tee $O/r2_train_phases_final.log
