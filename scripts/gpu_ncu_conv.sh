#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for mode in 1 2; do
  GLG_GEMM_CTA2=$mode timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_conv640_cta$mode -f \
     python scripts/bench_kernels.py x "conv3x3[${mode}cta] 64x64 640->640" > gpurun_out/ncu_conv_$mode.log 2>&1
  tail -n 2 gpurun_out/ncu_conv_$mode.log
done
GLG_GEMM_CTA2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_qkvL0_cta1 -f \
     python scripts/bench_kernels.py x "gemm[1cta] qkv                  M=32768" > gpurun_out/ncu_qkv.log 2>&1
tail -n 2 gpurun_out/ncu_qkv.log
