#!/bin/bash
# same-box A/B of library builds on the single-tower lines (BASELINE configs[3])
mkdir -p gpurun_out
L=gpurun_out/tower_ab.log
: > $L
for r in 1 2; do
  for lib in bridgelang_b200/libbridgelang_b200.so tools/_bin/lib_r2_base.so; do
    for t in siglip dino; do
      v=$(BLB_LIB=$lib timeout -k 5 100 python bench.py --tower $t --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['clocks']['sm_mhz'])")
      echo "$r $t $lib $v" >> $L
    done
  done
done
cat $L
