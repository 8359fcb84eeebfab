#!/bin/bash
# final round-1 evidence for the committed build: full GPU suite, default bench, smoke, launch list of the bench step,
# ncu full on the dominant kernel shape (each command first runs to completion WITHOUT ncu)
mkdir -p gpurun_out
L=gpurun_out/final_r01.log
: > $L
(timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4) >> $L
(timeout 600 python bench.py 2>&1 | tail -1) >> $L
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1) >> $L
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/plain_bench_d.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 300 --csv --log-file gpurun_out/launches_bench_d.csv $B > gpurun_out/ncu_bench_d.log 2>&1
timeout 200 python tools/bringup.py gemm_fold 2 65536 4352 1152 1 > gpurun_out/plain_d_fc1s.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/prof_d_fc1s_ln -f python tools/bringup.py gemm_fold 2 65536 4352 1152 1 > gpurun_out/ncu_d_fc1s.log 2>&1
timeout 200 python tools/bringup.py attn > gpurun_out/plain_d_attn.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tail -s 12 -c 1 -o gpurun_out/prof_d_attn_tail -f python tools/bringup.py attn > gpurun_out/ncu_d_attn_tail.log 2>&1
cat $L
