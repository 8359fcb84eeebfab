#!/bin/bash
# round-1 profiling pass: full GPU test-suite, ncu full capture of the dominant GEMM + attention, launch list of bench
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -40) > gpurun_out/pytest_gpu.log
G="python tools/bringup.py gemm 2 66816 3072 1024 0"
$G > gpurun_out/plain_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 2 -o gpurun_out/prof_gemm_qkv -f $G > gpurun_out/ncu_gemm.log 2>&1
G2="python tools/bringup.py gemm 2 66816 4096 1024 1"
$G2 > gpurun_out/plain_gemm2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/prof_gemm_fc1 -f $G2 > gpurun_out/ncu_gemm2.log 2>&1
A="python tools/bringup.py attn"
$A > gpurun_out/plain_attn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention -s 4 -c 2 -o gpurun_out/prof_attn -f $A > gpurun_out/ncu_attn.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/pytest_gpu.log; tail -3 gpurun_out/ncu_gemm.log gpurun_out/ncu_attn.log gpurun_out/ncu_bench.log; ls -la gpurun_out
