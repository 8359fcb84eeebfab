#!/bin/bash
# ncu --set full (+ source counters) of the tcgen05 attention kernel at B=256: DINOv2 (261 x 64) and SigLIP (256 x 72)
mkdir -p gpurun_out
timeout -k 5 120 python tools/bringup.py attn > gpurun_out/plain_attn.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 30 -c 1 -o gpurun_out/${1}_attn_dino -f python tools/bringup.py attn > gpurun_out/ncu_attn_dino.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 44 -c 1 -o gpurun_out/${1}_attn_siglip -f python tools/bringup.py attn > gpurun_out/ncu_attn_siglip.log 2>&1
tail -3 gpurun_out/plain_attn.log
