#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== debug f1b"; timeout 600 python tools/debug_f1b.py > $O/r2_debug_f1b.log 2>&1; tail -45 $O/r2_debug_f1b.log
echo "== ragan"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -s -m gpu -k "train_steps and ragan" 2>&1 | grep -v "^$" | tail -60
