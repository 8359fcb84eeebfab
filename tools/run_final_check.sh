#!/bin/bash
# last check of the committed tree: full GPU suite, default bench line, smoke
mkdir -p gpurun_out
L=gpurun_out/final_check_r02.log
: > $L
(timeout -k 5 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -3) >> $L
(timeout -k 5 300 python bench.py 2>&1 | tail -1) > gpurun_out/r02_bench_final2.json
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_final2.json').read())
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],2), 'attn_ms', round(d['roofline']['breakdown_ms_per_step']['attention'],2), 'frac', round(d['roofline']['frac'],3), 'burst', round(d['roofline']['whole_step']['frac_of_burst'],3), d['clocks'])" >> $L
(timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1) >> $L
cat $L
