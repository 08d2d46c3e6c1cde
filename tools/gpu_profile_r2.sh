#!/bin/bash
# Final profiling pass of round 2 (run through gpurun, ONE GPU).  Launch list of one run of the hot path, `ncu --set full`
# captures of one steady-state round, of the dominant engine kernel (with source), of the dense kernel-1 benchmark and of
# the offline kernels; per-category CUDA-event timing of the same command.  .ncu-rep files are exported to CSV / text on the
# box and dropped (size limit); tools/summarize_ncu.py turns the CSVs into the markdown files under profiles/.
tag=${1:-r2z}
out=gpurun_out
mkdir -p $out
timeout 300 python tools/quick_perf.py C2 1.0 > $out/${tag}_quick_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $out/${tag}_launches.csv \
    python tools/quick_perf.py C2 0.2 > $out/${tag}_ncu_launches.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none -k 'regex:k_bs_|k_nearest|k_topk' -c 40 -f \
    -o /tmp/${tag}_steady python tools/profile_window.py C2 1.0 100000 > $out/${tag}_ncu_steady.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_bs_chain_p -c 1 -f \
    -o /tmp/${tag}_chain_p python tools/profile_window.py C2 1.0 100000 > $out/${tag}_ncu_chain_p.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_nearest -c 1 -f \
    -o /tmp/${tag}_k1 python tools/profile_window.py k1 > $out/${tag}_ncu_k1.log 2>&1
timeout 300 ncu --set full --clock-control none -k 'regex:k_off_neighbours|k_off_weighted|k_off_subspace|k_offc_grow' -c 4 -f \
    -o /tmp/${tag}_off python tools/bench_offline.py --M 100000 --D 40 --reps 0 > $out/${tag}_ncu_off.log 2>&1
for r in steady chain_p k1 off; do
  if [ -f /tmp/${tag}_$r.ncu-rep ]; then
    ncu -i /tmp/${tag}_$r.ncu-rep --page raw --csv > $out/${tag}_${r}_raw.csv 2>/dev/null
    ncu -i /tmp/${tag}_$r.ncu-rep --page details > $out/${tag}_${r}_details.txt 2>/dev/null
  fi
done
ncu -i /tmp/${tag}_chain_p.ncu-rep --page source --csv > $out/${tag}_chain_p_source.csv 2>/dev/null
du -sh $out
