#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_tests.sh --quick > gpurun_out/tests_quick.txt 2>&1
cat gpurun_out/summary.txt
if grep -q "failed" gpurun_out/summary.txt; then
  echo "FAILURES with PDL -> retry with GLG_PDL=0"
  export GLG_PDL=0
  bash scripts/gpu_tests.sh --quick > gpurun_out/tests_quick_nopdl.txt 2>&1
  cat gpurun_out/summary.txt
fi
timeout 300 python scripts/bench_kernels.py r1g "gemm[auto]" 2>&1 | tee gpurun_out/kernels_r1g.txt | head -40
timeout 300 python scripts/bench_kernels.py r1g2 "groupnorm" 2>&1 | tee -a gpurun_out/kernels_r1g.txt | head -40
bash scripts/gpu_bench.sh r1g
GLG_PDL=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-pass > gpurun_out/bench_r1g_nopdl.json 2>&1
python -c "import json;d=json.load(open('gpurun_out/bench_r1g_nopdl.json'));print('no-PDL value',d['value'])"
