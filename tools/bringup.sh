#!/bin/bash
# runs every bring-up case in its own process under a timeout; output → gpurun_out/bringup.log
mkdir -p gpurun_out
L=gpurun_out/bringup.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $L 2>&1
run() { echo "### $*" >> $L; timeout 120 python tools/bringup.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run im2col
run argmax
run ln
run attn
for ctas in 1 2; do
  run gemm $ctas 256 256 128 0
  run gemm $ctas 384 512 1024 0
  run gemm $ctas 1000 384 1152 1
  run gemm $ctas 1000 768 592 2
  run gemm $ctas 66816 3072 1024 0
  run gemm $ctas 66816 4096 1024 1
  run gemm $ctas 66816 1024 4096 2
  run gemm $ctas 65536 3456 1152 0
  run gemm $ctas 65536 4352 1152 1
  run gemm $ctas 65536 1152 4352 2
  run gemm $ctas 65536 8704 2176 1
  run gemm $ctas 65536 4096 8704 1
done
tail -100 $L
