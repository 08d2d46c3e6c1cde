#!/bin/bash
tag=${1:-r2d}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
tail -4 $out/${tag}_pytest_gpu.log
for c in 0 49152 65536 131072; do
  timeout 300 python tools/tp_wall.py C2 1.0 $c > $out/${tag}_tp_wall_c2_chunk$c.log 2>&1; echo "chunk $c"; tail -1 $out/${tag}_tp_wall_c2_chunk$c.log
done
timeout 300 python tools/quick_perf.py C2 1.0 0 0 0 --chain > $out/${tag}_quick_c2.log 2>&1; grep -B3 -A3 "chain_p last" $out/${tag}_quick_c2.log | cut -c1-700
