#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run14.log
: > $L
for v in 2048 1024 2048 1024; do
  echo "## BLB_RESID_DEEPK=$v" >> $L
  (BLB_RESID_DEEPK=$v timeout 300 python tools/bringup.py gemm_fold 2 66816 1024 1024 2 2>&1 | tail -1) >> $L
  (BLB_RESID_DEEPK=$v timeout 300 python tools/bringup.py gemm_fold 2 65536 1152 1152 2 2>&1 | tail -1) >> $L
done
B="timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
run() { echo "## bench $1" >> $L; (env $1 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['breakdown_ms_per_step'])") >> $L 2>&1; }
run "BLB_RESID_DEEPK=2048"
run "BLB_RESID_DEEPK=1024"
run "BLB_RESID_DEEPK=2048"
run "BLB_RESID_DEEPK=1024"
cat $L
