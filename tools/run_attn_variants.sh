#!/bin/bash
# time attention-kernel build variants (tools/_bin/lib_attn_*.so) on one box; parity first
mkdir -p gpurun_out
L=gpurun_out/attn_variants.log
: > $L
for v in "$@"; do
  echo "== variant $v" >> $L
  (BLB_LIB=tools/_bin/lib_attn_$v.so timeout -k 5 90 python -m pytest tests/test_gpu_kernels.py -q -x -k attention --timeout 120 2>&1 | tail -2) >> $L
  (BLB_LIB=tools/_bin/lib_attn_$v.so timeout -k 5 90 python tools/bringup.py attn 2>&1 | grep -A1 "B=4 \|time" | grep -v "^--" | tail -8) >> $L
done
echo "== base" >> $L
(BLB_LIB=tools/_bin/lib_r2_base.so timeout -k 5 90 python tools/bringup.py attn 2>&1 | grep time | tail -2) >> $L
(BLB_LIB=tools/_bin/lib_attn_$1.so timeout -k 5 90 python tools/trace_attn.py > gpurun_out/attn_trace_$1.log 2>&1)
cat $L
