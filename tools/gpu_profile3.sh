#!/bin/bash
# Final profiling pass of a round (run through gpurun).  Steady-state ncu captures of the dominant engine kernel, the dense
# kernel-1 benchmark and the offline neighbourhood kernel, the launch list of one run of the hot path, and the same command's
# per-category CUDA-event timing.  .ncu-rep files are exported to CSV / text on the box and dropped (size limit).
tag=${1:-r1x}
out=gpurun_out
mkdir -p $out
timeout 300 python tools/quick_perf.py C2 1.0 0 0 0 --chain > $out/${tag}_quick_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file $out/${tag}_launches.csv \
    python tools/quick_perf.py C2 0.2 > $out/${tag}_ncu_launches.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_bs_chain_p -c 1 -f \
    -o /tmp/${tag}_chain_p python tools/profile_window.py C2 1.0 100000 > $out/${tag}_ncu_chain_p.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_nearest -c 1 -f \
    -o /tmp/${tag}_k1 python tools/profile_window.py k1 > $out/${tag}_ncu_k1.log 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none -k 'regex:k_bs_|k_nearest|k_topk' -c 40 -f \
    -o /tmp/${tag}_steady python tools/profile_window.py C2 1.0 100000 > $out/${tag}_ncu_steady.log 2>&1
timeout 300 ncu --set full --clock-control none -k 'regex:k_off_neighbours|k_off_weighted|k_off_subspace|k_offc_grow' -c 4 -f \
    -o /tmp/${tag}_off python tools/bench_offline.py --M 100000 --D 40 --reps 0 > $out/${tag}_ncu_off.log 2>&1
for r in chain_p k1 steady off; do
  if [ -f /tmp/${tag}_$r.ncu-rep ]; then
    ncu -i /tmp/${tag}_$r.ncu-rep --page raw --csv > $out/${tag}_${r}_raw.csv 2>/dev/null
    ncu -i /tmp/${tag}_$r.ncu-rep --page details > $out/${tag}_${r}_details.txt 2>/dev/null
  fi
done
ncu -i /tmp/${tag}_chain_p.ncu-rep --page source --csv > $out/${tag}_chain_p_source.csv 2>/dev/null
ncu -i /tmp/${tag}_k1.ncu-rep --page source --csv > $out/${tag}_k1_source.csv 2>/dev/null
# does ncu see the kernels inside the conditional-node graph?  (informational; the launch list above uses stream launches)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches_graph.csv \
    python tools/tp_wall.py C2 0.1 > $out/${tag}_ncu_launches_graph.log 2>&1
du -sh $out
