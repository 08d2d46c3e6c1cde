#!/bin/bash
# Steady-state ncu captures (run through gpurun): chain_p inside timepoint 2 of C2, and the dense kernel-1 benchmark.
# The .ncu-rep files (tens of MB with --import-source) are exported to CSV on the box and dropped.
tag=${1:-r1x}
out=gpurun_out
mkdir -p $out
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_bs_chain_p -c 2 -f \
    -o /tmp/${tag}_chain_p python tools/profile_window.py C2 0.5 100000 > $out/${tag}_ncu_chain_p.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_nearest -c 1 -f \
    -o /tmp/${tag}_k1 python tools/profile_window.py k1 > $out/${tag}_ncu_k1.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none -k 'regex:k_bs_|k_nearest|k_topk' -c 60 -f \
    -o /tmp/${tag}_steady python tools/profile_window.py C2 0.5 100000 > $out/${tag}_ncu_steady.log 2>&1
for r in chain_p k1 steady; do
  if [ -f /tmp/${tag}_$r.ncu-rep ]; then
    ncu -i /tmp/${tag}_$r.ncu-rep --page raw --csv > $out/${tag}_${r}_raw.csv 2>/dev/null
  fi
done
ncu -i /tmp/${tag}_chain_p.ncu-rep --page source --csv > $out/${tag}_chain_p_source.csv 2>/dev/null
ncu -i /tmp/${tag}_k1.ncu-rep --page source --csv > $out/${tag}_k1_source.csv 2>/dev/null
ncu -i /tmp/${tag}_k1.ncu-rep --page details > $out/${tag}_k1_details.txt 2>/dev/null
ncu -i /tmp/${tag}_chain_p.ncu-rep --page details > $out/${tag}_chain_p_details.txt 2>/dev/null
du -sh $out
