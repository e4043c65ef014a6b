#!/bin/bash
# C host replay test + selective programmatic dependent launch (GLG_PDL=2: only grids that leave SMs idle) against the default
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_native_engine_gpu.py -q -m gpu -s -k "c_host" 2>&1 | tail -4
for m in 0 2 0 2; do
  GLG_PDL=$m timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-pass > gpurun_out/bench_pdl$m.json 2> gpurun_out/bench_pdl$m.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_pdl$m.json')); print('GLG_PDL=$m', round(d['value'],4), 'img/s', round(d['ms_per_step'],2), 'ms', d['clocks']['reasons'])"
done
