#!/bin/bash
# Run the GPU parity suite group by group (separate processes: a sticky CUDA error in one group must not
# poison the others); logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
run() {
  name=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout 1500 python -m pytest "$@" -q -m gpu --timeout 1200 -s > gpurun_out/$name.log 2>&1
  echo "exit $?" >> gpurun_out/$name.log
  tail -n 3 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run small   tests/test_kernels_gpu.py -k "small_ops or groupnorm or layernorm"
run gemm    tests/test_kernels_gpu.py -k "gemm" --maxfail=12
run conv    tests/test_kernels_gpu.py -k "conv3x3" --maxfail=12
run attn    tests/test_kernels_gpu.py -k "attention" --maxfail=12
run tiny    tests/test_engine_gpu.py -k "tiny or scale_zero" --maxfail=12
if [ "$1" != "--quick" ]; then
  run sd14  tests/test_engine_gpu.py -k "sd14"
fi
grep -h "rel_l2" gpurun_out/tiny.log gpurun_out/sd14.log 2>/dev/null | tail -n 60 > gpurun_out/parity_numbers.txt
cat gpurun_out/summary.txt
