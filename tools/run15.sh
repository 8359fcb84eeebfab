#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run15.log
: > $L
(timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -5) >> $L
(timeout 600 python bench.py 2>&1 | tail -1) >> $L
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) >> $L
cat $L
