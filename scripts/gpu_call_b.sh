#!/bin/bash
# end-of-round confirmation on one B200: ncu --set full of the three-tile attention kernel, the whole -m gpu suite (as the driver runs it),
# smoke(), the default bench line (with the reference PLMSSampler on the host cores beside it), the reference arm, the config sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --plms-steps 1 --no-cpu-baseline --no-kernel-pass"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc3_kernel -s 3 -c 1 -o gpurun_out/prof_attn3_r2b -f $B > gpurun_out/ncu_attn3_r2b.log 2>&1; tail -n 1 gpurun_out/ncu_attn3_r2b.log | cut -c1-200
bash scripts/gpu_final.sh
timeout 1200 python bench.py --sweep --no-cpu-baseline > gpurun_out/bench_sweep_final.jsonl 2> gpurun_out/bench_sweep_final.err; python - <<'PY'
import json
for l in open('gpurun_out/bench_sweep_final.jsonl'):
    try:
        d = json.loads(l); print(round(d['value'], 3), d['unit'], d['ms_per_step'], d['config']['workload'][:110])
    except Exception as e:
        print('?', l[:100])
PY
tail -n 3 gpurun_out/bench_sweep_final.err
