#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --clock-control none -k regex:'conv2d_in_lrelu|conv2d_f32_kernel|conv2d_wgrad_f32|bn_partial|bn_apply|bias_grad_partial_bf16x8|instnorm_lrelu_bwd' -s 20 -c 36 -o /tmp/r2_side_kernels python tools/side_kernels.py > $O/r2_side_kernels.log 2>&1; tail -2 $O/r2_side_kernels.log
ncu -i /tmp/r2_side_kernels.ncu-rep --page raw --csv > /tmp/side_raw.csv 2>/dev/null
python tools/ncu_metrics.py /tmp/side_raw.csv > $O/r2_side_kernels_metrics.txt; wc -l $O/r2_side_kernels_metrics.txt
ls -la /tmp/r2_side_kernels.ncu-rep; SZ=$(stat -c %s /tmp/r2_side_kernels.ncu-rep); if [ "$SZ" -lt 40000000 ]; then cp /tmp/r2_side_kernels.ncu-rep $O/; fi
