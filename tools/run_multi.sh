#!/bin/bash
# multi-GPU check exactly as the driver launches it: N ranks, one per GPU, NCCL
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/run_multi_$N.log
: > $L
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
(timeout 600 $T bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep '^{' | tail -1) >> $L
(timeout 600 $T bench.py --gpus $N --steps 5 --warmup 3 --gather 2>&1 | grep '^{' | tail -1) >> $L
(timeout 600 $T bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | grep '^{' | tail -1) >> $L
cat $L
