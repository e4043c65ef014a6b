#!/bin/bash
# End-of-round confirmation: full GPU suite, smoke(), default bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_tests.sh > /dev/null 2>&1
cat gpurun_out/summary.txt | grep -E "===|passed|failed|error" 
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 1500 gpurun_out/bench_final.json; tail -n 3 gpurun_out/bench_final.err
