#!/bin/bash
# GPU parity of the rows added late in round 2 (spatial modalities, CLIP text encoder) + the attention suite, the three-tile
# attention kernel's variants, front-end timings and a bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spatial_gpu.py tests/test_clip_gpu.py tests/test_pipeline_gpu.py tests/test_native_engine_gpu.py -q -m gpu -s 2>&1 | tail -80 > gpurun_out/new_rows_gpu.txt; cat gpurun_out/new_rows_gpu.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" 2>&1 | tail -3
(timeout 200 env VARS=1 python scripts/micro/attn_l0.py 5 > gpurun_out/attn_l0_b.txt 2>&1); grep -E "variant|poly 2/8 stagger    0" gpurun_out/attn_l0_b.txt
(timeout 300 python scripts/bench_frontend.py 4 > gpurun_out/frontend_b4.jsonl 2> gpurun_out/frontend_b4.err); cat gpurun_out/frontend_b4.jsonl; tail -n 3 gpurun_out/frontend_b4.err
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_s2b.json 2> gpurun_out/bench_s2b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_s2b.json')); print('VALUE', d['value'], d['ms_per_step'], {k:(v['ms'],v['tflops']) for k,v in d['kernel_shares'].items() if v['share']>0.01})"; tail -n 2 gpurun_out/bench_s2b.err
