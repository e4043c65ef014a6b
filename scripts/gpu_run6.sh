#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pair kernel tests"
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 300 -k "gemm and 2-" 2>&1 | tail -n 15 | tee gpurun_out/pair_tests.txt
if grep -q "failed\|error" gpurun_out/pair_tests.txt; then
  echo "PAIR KERNEL BROKEN -> falling back to GLG_GEMM_CTA2=1 for the rest"
  export GLG_GEMM_CTA2=1
fi
bash scripts/gpu_tests.sh --quick > gpurun_out/tests_quick.txt 2>&1
cat gpurun_out/summary.txt
timeout 500 python scripts/bench_kernels.py r1e 2>&1 | tee gpurun_out/kernels_r1e.txt | grep -E "gemm|conv3x3" | head -120
bash scripts/gpu_bench.sh r1e
