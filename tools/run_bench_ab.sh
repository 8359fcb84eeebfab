#!/bin/bash
# interleaved A/B of library builds on one box: tools/run_bench_ab.sh <rounds> <lib> <lib> ...   (img/s, device-timed)
mkdir -p gpurun_out
L=gpurun_out/bench_ab.log
: > $L
R=$1; shift
for r in $(seq 1 $R); do
  for lib in "$@"; do
    v=$(BLB_LIB=$lib timeout -k 5 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], round(d['roofline']['breakdown_ms_per_step']['attention'],2))")
    echo "$r $lib $v" >> $L
  done
done
cat $L
