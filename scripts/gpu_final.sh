#!/bin/bash
# End-of-round confirmation: full GPU suite, smoke(), default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# exactly what the driver runs first: the whole -m gpu suite in one process
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1; tail -n 3 gpurun_out/pytest_gpu_all.log
timeout 600 python -m pytest tests/test_vae_gpu.py tests/test_batch_gpu.py -q -m gpu -s 2>&1 | grep -E "VAE|row|passed|failed" | tail -n 14 > gpurun_out/vae_batch.txt; tail -n 3 gpurun_out/vae_batch.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 1500 gpurun_out/bench_final.json; tail -n 3 gpurun_out/bench_final.err
# the driver's reference arm (same flags as the product arm)
timeout 1500 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 900 gpurun_out/bench_reference.json; tail -n 2 gpurun_out/bench_reference.err
