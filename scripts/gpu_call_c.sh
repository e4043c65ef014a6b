#!/bin/bash
# 2 GPUs of one box: shard-equivalence test (NCCL), BASELINE config 4 (inpaint, 16 images over 2 GPUs), config 2 weak and config 5 strong scaling
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_gpu.py -q -m gpu -s 2>&1 | tail -6 > gpurun_out/dist_2gpu_final.log; cat gpurun_out/dist_2gpu_final.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 2 --preset 4 --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-pass > gpurun_out/bench_config4_2gpu_final.json 2> gpurun_out/bench_c4.err; tail -c 400 gpurun_out/bench_config4_2gpu_final.json | head -c 10; python -c "
import json; d=json.load(open('gpurun_out/bench_config4_2gpu_final.json')); print('config4 2gpu', d['value'], d['ms_per_step'], d['n_gpus'])"
timeout 600 $T bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-pass > gpurun_out/bench_config2_2gpu_final.json 2> gpurun_out/bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_config2_2gpu_final.json')); print('config2 2gpu', d['value'], d['ms_per_step'], d['n_gpus'])"
timeout 600 $T bench.py --gpus 2 --preset 5 --global-batch 16 --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-pass > gpurun_out/bench_config5_strong_2gpu_final.json 2> gpurun_out/bench_c5.err; python -c "
import json; d=json.load(open('gpurun_out/bench_config5_strong_2gpu_final.json')); print('config5 strong 2gpu', d['value'], d['ms_per_step'], d['scaling'], d['config']['workload'][:90])"
tail -n 2 gpurun_out/bench_c4.err gpurun_out/bench_c5.err
