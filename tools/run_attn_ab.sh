#!/bin/bash
# A/B of the attention kernel on one box: bounded-wait build first (a protocol bug traps instead of hanging)
mkdir -p gpurun_out
L=gpurun_out/attn_ab.log
: > $L
echo "== bounded-wait build: parity" >> $L
(BLB_LIB=tools/_bin/lib_bounded.so timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k attention --timeout 120 2>&1 | tail -5) >> $L
echo "== new build: parity + timing" >> $L
(timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k attention --timeout 120 2>&1 | tail -3) >> $L
(timeout 300 python tools/bringup.py attn 2>&1 | tail -12) >> $L
echo "== base build: timing" >> $L
(BLB_LIB=tools/_bin/lib_r2_base.so timeout 300 python tools/bringup.py attn 2>&1 | tail -12) >> $L
echo "== new build: timing again" >> $L
(timeout 300 python tools/bringup.py attn 2>&1 | tail -12) >> $L
(timeout 300 python tools/trace_attn.py > gpurun_out/attn_trace_new.log 2>&1)
cat $L
