#!/bin/bash
# One profiling pass on the GPU box (run through gpurun): per-category event timing, ncu launch list, ncu --set full
# captures exported as CSV on the box (the .ncu-rep files are kept only while gpurun_out/ stays under the 64 MiB limit).
# usage: tools/gpu_profile.sh <tag>
tag=${1:-r1x}
out=gpurun_out
mkdir -p $out
timeout 300 python tools/quick_perf.py C2 1.0 0 0 0 --chain > $out/${tag}_quick_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_launches.csv \
    python tools/quick_perf.py C2 0.2 > $out/${tag}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bs_chain_p -s 30 -c 2 -f -o $out/${tag}_chain_p \
    python tools/quick_perf.py C2 0.2 > $out/${tag}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none -k 'regex:k_bs_(spec|need|tilecnt|pscan|pscatter|olist|chain_o|derive|verify_p|verify_o|decide|commit_rows|commit_cells|finish)|k_nearest|k_topk' \
    -s 400 -c 36 -f -o $out/${tag}_round python tools/quick_perf.py C2 0.2 > $out/${tag}_ncu3.log 2>&1
for r in chain_p round; do
  if [ -f $out/${tag}_$r.ncu-rep ]; then
    ncu -i $out/${tag}_$r.ncu-rep --page raw --csv > $out/${tag}_${r}_raw.csv 2>/dev/null
  fi
done
ncu -i $out/${tag}_chain_p.ncu-rep --page source --csv > $out/${tag}_chain_p_source.csv 2>/dev/null
rm -f $out/${tag}_round.ncu-rep
sz=$(du -sm $out | cut -f1)
if [ "$sz" -gt 50 ]; then rm -f $out/*.ncu-rep; fi
du -sh $out
