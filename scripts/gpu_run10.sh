#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout 1200 python -m pytest "$@" -q -m gpu --timeout 900 -s > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log
  tail -n 3 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
run small tests/test_kernels_gpu.py -k "small_ops or groupnorm or layernorm"
run attn  tests/test_kernels_gpu.py -k "attention" --maxfail=12
run tiny  tests/test_engine_gpu.py -k "tiny or scale_zero" --maxfail=12
timeout 600 python scripts/bench_kernels.py r1i attn > gpurun_out/kernels_r1i.txt 2>&1
timeout 600 python scripts/bench_kernels.py r1i2 groupnorm >> gpurun_out/kernels_r1i.txt 2>&1
cat gpurun_out/kernels_r1i.txt
bash scripts/gpu_bench.sh r1i
