#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout 1200 python -m pytest "$@" -q -m gpu --timeout 900 -s > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log
  tail -n 2 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
run small tests/test_kernels_gpu.py -k "small_ops or groupnorm or layernorm"
run attn  tests/test_kernels_gpu.py -k "attention"
run tiny  tests/test_engine_gpu.py -k "tiny or scale_zero" --maxfail=12
bash scripts/gpu_bench.sh r1n | head -3 | cut -c1-400
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1n.json'))
print({k:(round(v['ms'],3)) for k,v in d['kernel_shares'].items()})
PY
