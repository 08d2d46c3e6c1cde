#!/bin/bash
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
timeout 900 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_c2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"], d.get("e2e_pageable"), d.get("parity_checked",{}).get("equal"), d.get("c3"), d.get("serial_floor",{}).get("frac"), d.get("gpu_ms_per_category"))
PY
