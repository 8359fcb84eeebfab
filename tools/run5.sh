#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run5.log
: > $L
run() { echo "### $*" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run python tools/bringup.py attn
run python tools/trace_attn.py
for shape in "66816 3072 1024 0" "66816 1024 1024 2" "66816 4096 1024 1" "66816 1024 4096 2" "65536 3456 1152 0" "65536 1152 1152 2" "65536 4352 1152 1" "65536 1152 4352 2" "65536 8704 2176 1" "65536 4096 8704 1"; do
  run python tools/bringup.py gemm 2 $shape
done
(timeout 1200 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -8) >> $L
(timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -3) >> $L
cat $L
