#!/bin/bash
mkdir -p gpurun_out
A="python tools/bringup.py attn"
timeout 120 $A > gpurun_out/plain_attn4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 30 -c 1 -o gpurun_out/prof_attn_v3 -f $A > gpurun_out/ncu_attn4.log 2>&1
tail -3 gpurun_out/ncu_attn4.log
