#!/bin/bash
out=gpurun_out
mkdir -p $out
timeout 300 python tools/trace_rounds.py C2 1.0 --tps 2 --out $out/r2p_trace_c2.npz --detail 40 > $out/r2p_trace_c2.log 2>&1
timeout 300 python tools/trace_rounds.py C2 0.3 --tps 2 --eps 0.04 --out $out/r2p_trace_c2_eps004.npz --detail 20 > $out/r2p_trace_c2_eps004.log 2>&1
head -12 $out/r2p_trace_c2.log
