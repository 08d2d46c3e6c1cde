#!/bin/bash
tag=${1:-r2t}
out=gpurun_out
mkdir -p $out
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_nearest -c 1 -f \
    -o /tmp/${tag}_k1e python tools/profile_window.py C2 1.0 100000 > $out/${tag}_ncu_k1e.log 2>&1
ncu -i /tmp/${tag}_k1e.ncu-rep --page source --csv > $out/${tag}_k1e_source.csv 2>/dev/null
ncu -i /tmp/${tag}_k1e.ncu-rep --page details > $out/${tag}_k1e_details.txt 2>/dev/null
du -sh $out/${tag}_k1e_source.csv
