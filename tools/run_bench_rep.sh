#!/bin/bash
# repeat the default bench N times on one box: value / e2e (device, host ms per step) — run-to-run spread of the headline
mkdir -p gpurun_out
L=gpurun_out/bench_rep.log
: > $L
for r in $(seq 1 $1); do
  timeout -k 5 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print(round(d['value'],1), round(e['value'],1), round(e['device_ms_per_step'],2), round(e['host_ms_per_step'],2), d['clocks']['sm_mhz'])" >> $L
done
cat $L
