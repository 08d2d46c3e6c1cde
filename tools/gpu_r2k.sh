#!/bin/bash
tag=${1:-r2k}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "offline or cluster_growth or c1 or app_run" > $out/${tag}_pytest_offline.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_offline.log
tail -3 $out/${tag}_pytest_offline.log
timeout 300 python tools/bench_offline.py --M 100000 --D 40 > $out/${tag}_offline_c4_1gpu.json 2>&1; cat $out/${tag}_offline_c4_1gpu.json | cut -c1-900
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_offline_launches.csv python tools/bench_offline.py --M 100000 --D 40 --reps 0 > $out/${tag}_ncu_offline.log 2>&1
python tools/summarize_ncu.py launches $out/${tag}_offline_launches.csv "offline C4 launches" | head -30
