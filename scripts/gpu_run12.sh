#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_tests.sh > /dev/null 2>&1
cat gpurun_out/summary.txt
cat gpurun_out/parity_numbers.txt | tail -12
bash scripts/gpu_bench.sh r1k --profile 2>&1 | cut -c1-1500
