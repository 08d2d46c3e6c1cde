#!/bin/bash
# round 2, pass B: parity suite + per-timepoint wall + per-category timing + ncu of one steady-state round
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
tail -5 $out/${tag}_pytest_gpu.log
timeout 300 python tools/tp_wall.py C2 > $out/${tag}_tp_wall_c2_graph.log 2>&1; cat $out/${tag}_tp_wall_c2_graph.log
timeout 300 python tools/quick_perf.py C2 1.0 0 0 0 --chain > $out/${tag}_quick_c2.log 2>&1; grep -v "^      key [6-9]\|^      key [1-9][0-9]" $out/${tag}_quick_c2.log | cut -c1-900
timeout 300 ncu --profile-from-start off --set full --clock-control none -k 'regex:k_bs_|k_nearest|k_topk' -c 40 -f \
    -o /tmp/${tag}_steady python tools/profile_window.py C2 1.0 100000 > $out/${tag}_ncu_steady.log 2>&1
ncu -i /tmp/${tag}_steady.ncu-rep --page raw --csv > $out/${tag}_steady_raw.csv 2>/dev/null
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_bs_chain_p -c 1 -f \
    -o /tmp/${tag}_chain_p python tools/profile_window.py C2 1.0 100000 > $out/${tag}_ncu_chain_p.log 2>&1
ncu -i /tmp/${tag}_chain_p.ncu-rep --page source --csv > $out/${tag}_chain_p_source.csv 2>/dev/null
ncu -i /tmp/${tag}_chain_p.ncu-rep --page details > $out/${tag}_chain_p_details.txt 2>/dev/null
du -sh $out
