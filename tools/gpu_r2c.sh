#!/bin/bash
tag=${1:-r2c}
out=gpurun_out
mkdir -p $out
timeout 300 python tools/quick_perf.py C2 1.0 0 0 0 --chain > $out/${tag}_quick_c2.log 2>&1; grep -A6 "chain_p last" $out/${tag}_quick_c2.log | cut -c1-200
timeout 300 python tools/quick_perf.py C2 1.0 0 0 0 --chain --dbg=4 > $out/${tag}_quick_c2_nofence.log 2>&1; grep -A6 "chain_p last" $out/${tag}_quick_c2_nofence.log | tail -16 | cut -c1-200
