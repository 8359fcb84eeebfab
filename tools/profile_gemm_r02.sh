#!/bin/bash
# ncu --set full of the three GEMM shapes that matter most in the final round-2 step (each first runs without ncu)
mkdir -p gpurun_out
for spec in "fc2s:2 65536 1152 4352 2" "fc1s:2 65536 4352 1152 1" "proj:2 66816 1024 1024 2"; do
  name=${spec%%:*}; args=${spec#*:}
  timeout -k 5 120 python tools/bringup.py gemm_fold $args > gpurun_out/plain_r02_$name.log 2>&1 || continue
  ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/r02_prof_gemm_$name -f python tools/bringup.py gemm_fold $args > gpurun_out/ncu_r02_$name.log 2>&1
  tail -1 gpurun_out/plain_r02_$name.log
done
