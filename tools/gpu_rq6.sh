#!/bin/bash
# experiment rq6: adaptive (two / eight cells per step) CONTESTED groups against always-pairs and always-batch, short first
# block of a cold run; deviation counters from the debug build
tag=${1:-rq6}
out=gpurun_out
mkdir -p $out
VP=chronoclust_b200/libccb_variant_pairs.so
VB=chronoclust_b200/libccb_variant_batch.so
timeout 300 python tools/tp_wall.py C2 1.0 > $out/${tag}_tp_adaptive.log 2>&1; grep -E "^#|rep 2" $out/${tag}_tp_adaptive.log | cut -c1-420
timeout 300 python tools/tp_wall.py C2 1.0 --lib $VP > $out/${tag}_tp_pairs.log 2>&1; grep -E "^#|rep 2" $out/${tag}_tp_pairs.log | cut -c1-420
timeout 300 python tools/tp_wall.py C2 1.0 --lib $VB > $out/${tag}_tp_batch.log 2>&1; grep -E "^#|rep 2" $out/${tag}_tp_batch.log | cut -c1-420
timeout 300 python tools/tp_wall.py C2 0.3 --eps 0.04 --tps 2 --reps 2 > $out/${tag}_tp_eps004_adaptive.log 2>&1; grep -E "^#|rep 1" $out/${tag}_tp_eps004_adaptive.log | cut -c1-300
timeout 300 python tools/trace_rounds.py C2 1.0 --tps 2 --out $out/${tag}_trace_c2.npz --detail 0 > $out/${tag}_trace_c2.log 2>&1; head -3 $out/${tag}_trace_c2.log | cut -c1-300
timeout 300 python tools/trace_rounds.py C2 0.3 --tps 2 --eps 0.04 --out $out/${tag}_trace_c2_eps004.npz --detail 0 > $out/${tag}_trace_c2_eps004.log 2>&1; head -2 $out/${tag}_trace_c2_eps004.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --no-c3 --no-c4 > $out/${tag}_bench_quick.json 2> $out/${tag}_bench_quick.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_quick.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["e2e_pageable"]["value"], d.get("parity_checked",{}).get("equal"))
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stress or c1 or C5 or host or device_scaler" > $out/${tag}_pytest_subset.log 2>&1; tail -1 $out/${tag}_pytest_subset.log
