#!/bin/bash
# round-1 profiling pass after the LN fold: ncu full on the dominant GEMM shapes as the towers run them + attention,
# and the launch list of one bench step (every command first runs to completion WITHOUT ncu)
mkdir -p gpurun_out
prof() { # name, skip, count, kernel-regex, cmd...
  local name=$1 skip=$2 cnt=$3 rx=$4; shift 4
  timeout 200 "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o gpurun_out/prof_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
}
prof c_qkv_ln 2 1 gemm_bf16 python tools/bringup.py gemm_fold 2 66816 3072 1024 0
prof c_fc1_ln 2 1 gemm_bf16 python tools/bringup.py gemm_fold 2 66816 4096 1024 1
prof c_fc2_stats 2 1 gemm_bf16 python tools/bringup.py gemm_fold 2 66816 1024 4096 2
prof c_proj_stats 2 1 gemm_bf16 python tools/bringup.py gemm_fold 2 66816 1024 1024 2
prof c_fc1s_ln 2 1 gemm_bf16 python tools/bringup.py gemm_fold 2 65536 4352 1152 1
prof c_fc2s_stats 2 1 gemm_bf16 python tools/bringup.py gemm_fold 2 65536 1152 4352 2
prof c_attn_dino 30 1 attention_tc python tools/bringup.py attn
prof c_attn_tail 30 1 attention_tail python tools/bringup.py attn
prof c_attn_siglip 44 1 attention_tc python tools/bringup.py attn
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/plain_bench_c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 300 --csv --log-file gpurun_out/launches_bench_c.csv $B > gpurun_out/ncu_bench_c.log 2>&1
tail -2 gpurun_out/plain_bench_c.log
