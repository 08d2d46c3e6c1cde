#!/bin/bash
tag=${1:-r2o}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python tools/tp_wall.py C2 1.0 > $out/${tag}_tp_wall_c2.log 2>&1; tail -1 $out/${tag}_tp_wall_c2.log
timeout 300 ncu --profile-from-start off --set full --clock-control none -k 'regex:k_bs_|k_nearest|k_topk' -c 20 -f \
    -o /tmp/${tag}_steady python tools/profile_window.py C2 1.0 100000 > $out/${tag}_ncu_steady.log 2>&1
ncu -i /tmp/${tag}_steady.ncu-rep --page raw --csv > $out/${tag}_steady_raw.csv 2>/dev/null
python tools/summarize_ncu.py raw $out/${tag}_steady_raw.csv "steady" 20 | cut -d'|' -f2,3,6,7,8 
