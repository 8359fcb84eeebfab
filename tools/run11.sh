#!/bin/bash
# A/B of two library builds on the same box: libblb_prev.so (previous) vs the in-tree build
mkdir -p gpurun_out
L=gpurun_out/run11.log
: > $L
(timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_towers.py -m gpu -q --timeout 300 -x 2>&1 | tail -3) >> $L
B="timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
run() { echo "## $1" >> $L; (env $1 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])") >> $L 2>&1; }
P=$PWD/bridgelang_b200/libblb_prev.so
run "BLB_X=0"
run "BLB_LIB=$P"
run "BLB_X=0"
run "BLB_LIB=$P"
cat $L
