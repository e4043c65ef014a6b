#!/bin/bash
# GPU parity of the rows added late in round 2 (spatial modalities, CLIP text encoder) + the attention suite + a bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spatial_gpu.py tests/test_clip_gpu.py -q -m gpu -s 2>&1 | tail -80 > gpurun_out/new_rows_gpu.txt; cat gpurun_out/new_rows_gpu.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" 2>&1 | tail -3
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_s2b.json 2> gpurun_out/bench_s2b.err; tail -c 600 gpurun_out/bench_s2b.json | head -c 600; python -c "
import json; d=json.load(open('gpurun_out/bench_s2b.json')); print('\nVALUE', d['value'], d['ms_per_step'], {k:(v['ms'],v['tflops']) for k,v in d['kernel_shares'].items() if v['share']>0.01})"
