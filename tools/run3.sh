#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run3.log
: > $L
(timeout 120 python tools/debug_resid.py) >> $L 2>&1
A="python tools/bringup.py attn"
timeout 120 $A > gpurun_out/plain_attn2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 2 -o gpurun_out/prof_attn_tc -f $A > gpurun_out/ncu_attn2.log 2>&1
cat $L; tail -5 gpurun_out/ncu_attn2.log
