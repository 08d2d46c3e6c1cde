#!/bin/bash
# round 2, pass A: parity suite + smoke + per-timepoint wall + per-category timing (with the replay kernel's cycle counters)
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
tail -5 $out/${tag}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $out/${tag}_smoke.log 2>&1; tail -3 $out/${tag}_smoke.log
timeout 300 python tools/tp_wall.py C2 > $out/${tag}_tp_wall_c2_graph.log 2>&1; cat $out/${tag}_tp_wall_c2_graph.log
timeout 300 python tools/quick_perf.py C2 1.0 0 0 0 --chain > $out/${tag}_quick_c2.log 2>&1; grep -v "^      key" $out/${tag}_quick_c2.log | cut -c1-900
