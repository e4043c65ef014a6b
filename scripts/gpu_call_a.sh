#!/bin/bash
# round 2, late: attention busy-poll variants, new-row parity with full output, profile pass of the final code
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
(timeout 200 env VARS=1 VARLIST=2,10,18,26,0,2 python scripts/micro/attn_l0.py 5 > gpurun_out/attn_l0_c.txt 2>&1); grep -E "variant" gpurun_out/attn_l0_c.txt
timeout 900 python -m pytest tests/test_spatial_gpu.py tests/test_clip_gpu.py tests/test_pipeline_gpu.py tests/test_native_engine_gpu.py -q -m gpu -s > gpurun_out/new_rows_gpu_full.txt 2>&1
grep -E "rel-L2|rel_l2|passed|failed|chain" gpurun_out/new_rows_gpu_full.txt | grep -v Warning | tail -30
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "tc3_variants" 2>&1 | tail -2
bash scripts/gpu_r2_profile.sh r2b 2>&1 | tail -20
