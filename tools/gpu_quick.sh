#!/bin/bash
# quick experiment harness: parity of the benched configuration against the oracle + timings, a parity subset of the GPU
# suite, per-timepoint wall times, round timelines of C2 and of the saturated corner
tag=${1:-rq}
out=gpurun_out
mkdir -p $out
timeout 600 python bench.py --steps 3 --warmup 3 --no-c3 --no-c4 > $out/${tag}_bench_quick.json 2> $out/${tag}_bench_quick.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_quick.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d.get("parity_checked",{}).get("equal"))
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "baseline_sizes or stress or c1" > $out/${tag}_pytest_subset.log 2>&1; tail -1 $out/${tag}_pytest_subset.log
timeout 300 python tools/tp_wall.py C2 1.0 > $out/${tag}_tp_wall_c2.log 2>&1; tail -1 $out/${tag}_tp_wall_c2.log
timeout 300 python tools/trace_rounds.py C2 1.0 --tps 2 --out $out/${tag}_trace_c2.npz --detail 0 > $out/${tag}_trace_c2.log 2>&1
timeout 300 python tools/trace_rounds.py C2 0.3 --tps 2 --eps 0.04 --out $out/${tag}_trace_c2_eps004.npz --detail 0 > $out/${tag}_trace_c2_eps004.log 2>&1
grep "== t" $out/${tag}_trace_c2_eps004.log
