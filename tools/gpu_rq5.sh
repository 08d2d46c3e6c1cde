#!/bin/bash
# experiment rq5: batched CONTESTED groups (eight cells per pass, predicted verdicts) against the paired version, result
# drain behind the engine (e2e), engine knob sweep -- one short call
tag=${1:-rq5}
out=gpurun_out
mkdir -p $out
V=chronoclust_b200/libccb_variant_pairs.so
timeout 600 python bench.py --steps 3 --warmup 3 --no-c3 --no-c4 > $out/${tag}_bench_quick.json 2> $out/${tag}_bench_quick.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_quick.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["e2e_pageable"]["value"], d.get("parity_checked",{}).get("equal"))
PY
timeout 300 python tools/tp_wall.py C2 1.0 --sweep ";chunk=24576;chunk=49152;iters=12;iters=16;bmin=4096" > $out/${tag}_tp_sweep_batch.log 2>&1; grep -E "^#|rep 2" $out/${tag}_tp_sweep_batch.log | cut -c1-400
timeout 300 python tools/tp_wall.py C2 1.0 --lib $V > $out/${tag}_tp_pairs.log 2>&1; grep -E "^#|rep 2" $out/${tag}_tp_pairs.log | cut -c1-400
timeout 300 python tools/tp_wall.py C2 0.3 --eps 0.04 --tps 2 --reps 2 > $out/${tag}_tp_eps004_batch.log 2>&1; grep -E "^#|rep 1" $out/${tag}_tp_eps004_batch.log | cut -c1-300
timeout 300 python tools/tp_wall.py C2 0.3 --eps 0.04 --tps 2 --reps 2 --lib $V > $out/${tag}_tp_eps004_pairs.log 2>&1; grep -E "^#|rep 1" $out/${tag}_tp_eps004_pairs.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stress or c1 or C5 or host or device_scaler" > $out/${tag}_pytest_subset.log 2>&1; tail -1 $out/${tag}_pytest_subset.log
