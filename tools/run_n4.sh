#!/bin/bash
# BASELINE configs[2] at 4 GPUs (512 images per GPU) + the overlapped prefix gather at 256 per GPU, launched as the driver launches them
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
(timeout -k 5 200 $T bench.py --gpus 4 --steps 5 --warmup 3 --global-batch 2048 --no-cpu-baseline 2>&1 | grep "^{" | tail -1) > gpurun_out/r02_bench_n4_g2048.json
(timeout -k 5 150 $T bench.py --gpus 4 --steps 5 --warmup 3 --gather --no-cpu-baseline 2>&1 | grep "^{" | tail -1) > gpurun_out/r02_bench_n4_gather_p2p.json
(timeout -k 5 150 $T bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | tail -1) > gpurun_out/r02_bench_n4.json
for f in r02_bench_n4_g2048 r02_bench_n4_gather_p2p r02_bench_n4; do python -c "
import json,sys
d=json.loads(open('gpurun_out/$f.json').read()); print('$f', round(d['value'],1), d['ms_per_step'], d['config']['batch_per_gpu'], d['config']['collective'][:60])"; done
