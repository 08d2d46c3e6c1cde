#!/bin/bash
tag=${1:-r2l}
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gating" 2>&1 | tail -2
for th in 1.0 1.5 2.0 3.0 4.0 8.0 1e9; do
  export CCB_THETA=$th
  timeout 300 python tools/tp_wall.py C2 1.0 --debuglib > $out/${tag}_tp_wall_c2_theta$th.log 2>&1; echo "theta $th"; tail -1 $out/${tag}_tp_wall_c2_theta$th.log
done
