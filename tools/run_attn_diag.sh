#!/bin/bash
# timing only (diagnostic builds have wrong numerics)
mkdir -p gpurun_out
L=gpurun_out/attn_diag.log
: > $L
for v in "$@"; do
  echo "== variant $v" >> $L
  (BLB_LIB=tools/_bin/lib_attn_$v.so timeout -k 5 90 python tools/bringup.py attn 2>&1 | grep time | tail -2) >> $L
done
cat $L
