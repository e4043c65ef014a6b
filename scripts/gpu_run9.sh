#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_tests.sh --quick > gpurun_out/tests_quick.txt 2>&1
cat gpurun_out/summary.txt
grep -h "FAILED\|^E  " gpurun_out/*.log | head -20
timeout 300 python scripts/bench_kernels.py r1h "attn" 2>&1 | tee gpurun_out/kernels_r1h.txt | head -40
timeout 300 python scripts/bench_kernels.py r1h2 "conv3x3[1cta] 8x8" 2>&1 | tee -a gpurun_out/kernels_r1h.txt | head
timeout 300 python scripts/bench_kernels.py r1h3 "M=512" 2>&1 | grep auto | tee -a gpurun_out/kernels_r1h.txt | head
bash scripts/gpu_bench.sh r1h
