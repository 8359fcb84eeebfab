#!/bin/bash
# validation pass: GPU tests, the driver's default bench invocation (with cpu_baseline), the reference arm
mkdir -p gpurun_out
L=gpurun_out/run8.log
: > $L
(timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8) >> $L
(time timeout 600 python bench.py 2>&1 | tail -1) >> $L 2>&1
(time timeout 600 python bench.py --impl reference 2>&1 | tail -1) >> $L 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) >> $L
cat $L
