#!/bin/bash
# LN-fold validation: GPU tests, bench folded vs explicit LayerNorm
mkdir -p gpurun_out
L=gpurun_out/run9.log
: > $L
(timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -15) >> $L
(timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) >> $L
(BLB_LN_EXPLICIT=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) >> $L
(timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) >> $L
cat $L
