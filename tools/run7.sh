#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run7.log
: > $L
(timeout 1200 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -4) >> $L
(timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) >> $L
(BLB_NO_PDL=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) >> $L
(timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) >> $L
cat $L
