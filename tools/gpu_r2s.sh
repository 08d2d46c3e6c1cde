#!/bin/bash
tag=${1:-r2s}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python tools/tp_wall.py C2 1.0 > $out/${tag}_tp_wall_c2.log 2>&1; tail -1 $out/${tag}_tp_wall_c2.log
timeout 300 python tools/tp_wall.py C3 0.5 > $out/${tag}_tp_wall_c3.log 2>&1; tail -1 $out/${tag}_tp_wall_c3.log
timeout 300 python tools/trace_rounds.py C2 1.0 --tps 2 --out $out/${tag}_trace_c2.npz --detail 0 > $out/${tag}_trace_c2.log 2>&1
timeout 300 python tools/trace_rounds.py C3 0.5 --tps 2 --out $out/${tag}_trace_c3.npz --detail 0 > $out/${tag}_trace_c3.log 2>&1
timeout 300 python tools/trace_rounds.py C2 0.3 --tps 2 --eps 0.04 --out $out/${tag}_trace_c2_eps004.npz --detail 0 > $out/${tag}_trace_c2_eps004.log 2>&1
head -1 $out/${tag}_trace_c2_eps004.log | cut -c1-200
