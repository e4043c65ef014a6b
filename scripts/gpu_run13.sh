#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout 1200 python -m pytest "$@" -q -m gpu --timeout 900 -s > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log
  tail -n 2 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
run tiny  tests/test_engine_gpu.py -k "tiny or scale_zero" --maxfail=12
run sd14  tests/test_engine_gpu.py -k "sd14"
grep -h "rel_l2" gpurun_out/sd14.log | tail -n 4
bash scripts/gpu_bench.sh r1m | head -3 | cut -c1-700
