#!/bin/bash
tag=${1:-r2x}
out=gpurun_out
mkdir -p $out
timeout 300 python tools/trace_rounds.py C2 1.0 --tps 1 --out $out/${tag}_trace_c2.npz --detail 0 > $out/${tag}_trace_c2.log 2>&1
grep "== t\|debug counters" $out/${tag}_trace_c2.log
