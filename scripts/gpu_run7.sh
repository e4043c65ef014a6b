#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_tests.sh --quick > gpurun_out/tests_quick.txt 2>&1
cat gpurun_out/summary.txt
timeout 500 python scripts/bench_kernels.py r1f "gemm" 2>&1 | tee gpurun_out/kernels_r1f.txt | grep -E "1cta|2cta" | head -80
timeout 300 python scripts/bench_kernels.py r1f2 "conv3x3" 2>&1 | tee -a gpurun_out/kernels_r1f.txt | head -40
bash scripts/gpu_bench.sh r1f
