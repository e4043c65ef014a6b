#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_tests.sh --quick > /dev/null 2>&1
cat gpurun_out/summary.txt
timeout 200 python scripts/micro/attn_probe.py 2>&1 | grep -E "s1 kernel|pair kernel|wait s_full" | cut -c1-220
timeout 900 python scripts/bench_kernels.py r1j > gpurun_out/kernels_r1j.txt 2>&1
grep -E "auto\]|\[auto|1cta\] square|conv3x3\[1cta\]|groupnorm" gpurun_out/kernels_r1j.txt | grep -v "attn cross"
bash scripts/gpu_bench.sh r1j | head -3 | cut -c1-1200
