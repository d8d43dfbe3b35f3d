#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== train step timing (skip-add fusion)"; TRAIN_PREC=bf16 STEPS=20 timeout 600 python tools/one_train_step.py 2>&1 | tail -2
echo "== parity of the mixed-precision step"; timeout 900 python -m pytest tests -q -s -m gpu -k "mixed or bf16 or train_steps or config1" 2>&1 | grep -n "passed\|failed\|FAILED\|mixed-prec\|config1" | head
echo "== small kernels under ncu"
DASR_B200_GRAPH=0 timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
   -k regex:'^(?!.*(conv_tc|wgrad_rdb|wgrad_tc|mma)).*$' --csv --log-file $O/r2_small_kernels.csv python tools/small_kernels.py > $O/r2_small_kernels.log 2>&1
tail -3 $O/r2_small_kernels.log; wc -l $O/r2_small_kernels.csv
python tools/summarize_small_kernels.py $O/r2_small_kernels.csv > $O/r2_small_kernels.md 2>&1; head -50 $O/r2_small_kernels.md
