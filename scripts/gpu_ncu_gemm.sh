#!/bin/bash
# ncu --set full (+ source page) of single launches: level-0 token GEMMs and the d_head = 40 attention variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for case in cxc qkvln ff1; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_r2_$case -f \
     python scripts/micro/gemm_one.py $case > gpurun_out/ncu_r2_$case.log 2>&1
  tail -n 1 gpurun_out/ncu_r2_$case.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc2 -s 2 -c 1 -o gpurun_out/prof_r2_attn_tc2 -f \
   python scripts/micro/attn_one.py 0 > gpurun_out/ncu_r2_attn_tc2.log 2>&1
tail -n 1 gpurun_out/ncu_r2_attn_tc2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_fused -s 2 -c 1 -o gpurun_out/prof_r2_gn -f \
   python scripts/bench_kernels.py x "groupnorm+silu HW=4096 C=320" > gpurun_out/ncu_r2_gn.log 2>&1
tail -n 1 gpurun_out/ncu_r2_gn.log
ls -la gpurun_out/*.ncu-rep
