#!/bin/bash
# 8-GPU pass (gpurun --gpus 8), trimmed and ordered by importance (the call may be cut by the GPU budget): bench.py at N = 8
# (replicas of C2 + the row-sharded offline stage C4), the 64-config parameter sweep C5 on 8 GPUs, then bench.py at N = 4, 2.
# (The N = 1 line is the one of the 1-GPU pass.)
tag=${1:-r2y}
out=gpurun_out
mkdir -p $out
bench() {
  N=$1
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 3 --warmup 3 --no-c3 > $out/${tag}_bench_c2_${N}gpu.json 2> $out/${tag}_bench_c2_${N}gpu.err; echo "bench $N rc=$?"
}
bench 8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 tools/sweep.py > $out/${tag}_sweep64_8gpu.json 2> $out/${tag}_sweep64_8gpu.err; echo "sweep rc=$?"
bench 4
bench 2
python - <<PY
import json
for N in (2,4,8):
    try:
        s=open("$out/${tag}_bench_c2_%dgpu.json"%N).read(); j=json.loads(s[s.index('{"metric'):].splitlines()[0])
        c=j["offline_c4"]
        print(N, "value %.1fM e2e %.1fM | C4 %.2f ms stages %s same=%s" % (j["value"]/1e6, j["e2e"]["value"]/1e6, c["ms"], c["stage_ms_rank0"], c["equals_1rank_result"]), j["per_rank_ms_per_step"], j["replicas_identical"])
    except Exception as e:
        print(N, "failed", e)
try:
    s=open("$out/${tag}_sweep64_8gpu.json").read(); j=json.loads(s[s.index('{"metric'):].splitlines()[0])
    print({k:j[k] for k in j if k!="runs"})
except Exception as e:
    print("sweep failed", e)
PY
