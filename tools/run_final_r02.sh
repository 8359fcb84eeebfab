#!/bin/bash
# round-2 evidence for the committed build: full GPU suite, default bench, smoke, launch list of the bench step
# (each command first runs to completion WITHOUT ncu), ncu full of the attention kernel at B=256
mkdir -p gpurun_out
L=gpurun_out/final_r02.log
: > $L
(timeout -k 5 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4) >> $L
(timeout -k 5 400 python bench.py 2>&1 | tail -1) > gpurun_out/r02_bench_final.json
cat gpurun_out/r02_bench_final.json >> $L
(timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1) >> $L
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout -k 5 300 $B > gpurun_out/plain_bench_r02.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 300 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/ncu_bench_r02.log 2>&1
bash tools/profile_attn.sh r02_final >> $L 2>&1
cat $L
