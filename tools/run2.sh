#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run2.log
: > $L
run() { echo "### $*" >> $L; timeout 120 "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run tools/_bin/tmem_bench
run python tools/bringup.py attn
BLB_ATTN_V1=1 run python tools/bringup.py attn
for shape in "66816 1024 1024 2" "65536 1152 1152 2" "66816 1024 4096 2" "65536 1152 4352 2" "66816 4096 1024 1" "65536 4352 1152 1" "66816 3072 1024 0"; do
  run python tools/bringup.py gemm 2 $shape
done
echo "### resid direct" >> $L
BLB_RESID_DIRECT=1 run python tools/bringup.py gemm 2 66816 1024 1024 2
(timeout 1200 python -m pytest tests -m gpu -q -x --timeout 180 2>&1 | tail -30) >> $L
(timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -3) >> $L
cat $L
