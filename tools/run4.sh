#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run4.log
: > $L
run() { echo "### $*" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run python tools/debug_resid.py
run python tools/bringup.py attn
for shape in "66816 1024 1024 2" "65536 1152 1152 2" "66816 1024 4096 2" "65536 1152 4352 2"; do
  run python tools/bringup.py gemm 2 $shape
done
(timeout 1200 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -15) >> $L
(timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -3) >> $L
A="python tools/bringup.py attn"
timeout 120 $A > gpurun_out/plain_attn3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 30 -c 1 -o gpurun_out/prof_attn_tc256 -f $A > gpurun_out/ncu_attn3.log 2>&1
cat $L; tail -3 gpurun_out/ncu_attn3.log
