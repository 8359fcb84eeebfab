#!/bin/bash
# round-1 final profiling pass: ncu full on the dominant GEMM shapes + attention, launch list of the bench step
mkdir -p gpurun_out
prof() { # name, skip, count, kernel-regex, cmd...
  local name=$1 skip=$2 cnt=$3 rx=$4; shift 4
  timeout 200 "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o gpurun_out/prof_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
}
prof gemm_qkv 2 1 gemm_bf16 python tools/bringup.py gemm 2 66816 3072 1024 0
prof gemm_fc1 2 1 gemm_bf16 python tools/bringup.py gemm 2 66816 4096 1024 1
prof gemm_fc2 2 1 gemm_bf16 python tools/bringup.py gemm 2 66816 1024 4096 2
prof gemm_proj 2 1 gemm_bf16 python tools/bringup.py gemm 2 66816 1024 1024 2
prof attn_dino 30 1 attention_tc python tools/bringup.py attn
prof attn_siglip 44 1 attention_tc python tools/bringup.py attn
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 748 -c 400 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/plain_bench.log
