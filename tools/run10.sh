#!/bin/bash
# A/B: tower overlap and deep-K residual staging (same box, back to back)
mkdir -p gpurun_out
L=gpurun_out/run10.log
: > $L
(timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -5) >> $L
B="timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
run() { echo "## $1" >> $L; (env $1 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])") >> $L 2>&1; }
run "BLB_X=0"
run "BLB_NO_TOWER_OVERLAP=1"
run "BLB_RESID_DEEPK=100000"
run "BLB_NO_TOWER_OVERLAP=1 BLB_RESID_DEEPK=100000"
run "BLB_X=0"
cat $L
