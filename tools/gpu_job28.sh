#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv2d_in_lrelu|conv2d_f32_kernel|conv2d_wgrad_f32|bn_partial|bn_apply|bn_finalize|bias_grad_partial_bf16x8|bias_grad_reduce|instnorm_lrelu_bwd' -s 20 -c 75 -o $O/r2_side_kernels python tools/side_kernels.py > $O/r2_side_kernels.log 2>&1; tail -2 $O/r2_side_kernels.log; ls -la $O/r2_side_kernels.ncu-rep
