#!/bin/bash
# the driver's N=2 launches with the final tree: native arm and reference arm
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
(timeout -k 5 200 $T bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | grep '^{' | tail -1) > gpurun_out/r02_bench_n2_final.json
(timeout -k 5 200 $T bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | grep '^{' | tail -1) > gpurun_out/r02_bench_n2_reference.json
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n2_final.json').read()); print('native', round(d['value'],1), round(d['e2e']['value'],1), d['n_gpus'], d['config']['parallelism'])
r=json.loads(open('gpurun_out/r02_bench_n2_reference.json').read()); print('reference', r.get('value'), r.get('impl'), r.get('n_gpus'))"
