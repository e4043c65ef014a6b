#!/bin/bash
# Run the GPU parity suite group by group (separate processes: a sticky CUDA error in one group must not
# poison the others); logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
run() {
  name=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout 1500 python -m pytest "$@" -q -m gpu --timeout 1200 -s --durations=8 > gpurun_out/$name.log 2>&1
  echo "exit $?" >> gpurun_out/$name.log
  grep -E "passed|failed|error" gpurun_out/$name.log | tail -n 2 | tee -a gpurun_out/summary.txt
  grep -E "^[0-9.]+s (call|setup)" gpurun_out/$name.log | head -n 4 | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run small   tests/test_kernels_gpu.py -k "small_ops or groupnorm or layernorm"
run gemm    tests/test_kernels_gpu.py -k "gemm" --maxfail=12
run conv    tests/test_kernels_gpu.py -k "conv3x3" --maxfail=12
run attn    tests/test_kernels_gpu.py -k "attention" --maxfail=12
run tiny    tests/test_engine_gpu.py -k "tiny or scale_zero" --maxfail=12
if [ "$1" != "--quick" ]; then
  run sd14  tests/test_engine_gpu.py -k "sd14"
  run batch tests/test_batch_gpu.py
  run final tests/test_final_latent_gpu.py
  run dist  tests/test_dist_gpu.py
fi
grep -h -E "rel_l2|rel-L2|FINAL LATENT|tolerance" gpurun_out/tiny.log gpurun_out/sd14.log gpurun_out/batch.log gpurun_out/final.log gpurun_out/dist.log 2>/dev/null | tail -n 120 > gpurun_out/parity_numbers.txt
cat gpurun_out/summary.txt
