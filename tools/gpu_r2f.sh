#!/bin/bash
# multi-GPU pass (gpurun --gpus N): NCCL parity test of the sharded offline phase, bench.py at N GPUs, offline C4 at N GPUs
N=${1:-2}
tag=${2:-r2f}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nccl" > $out/${tag}_pytest_nccl_${N}gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_nccl_${N}gpu.log
tail -4 $out/${tag}_pytest_nccl_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 5 --warmup 3 > $out/${tag}_bench_c2_${N}gpu.json 2> $out/${tag}_bench_c2_${N}gpu.err; echo "bench rc=$?"; tail -3 $out/${tag}_bench_c2_${N}gpu.err
python - <<PY
import json
j=json.load(open("$out/${tag}_bench_c2_${N}gpu.json"))
print({k:j[k] for k in ("value","n_gpus","ms_per_step","per_rank_ms_per_step","replicas_identical")}, j["e2e"]["value"], j["e2e_pageable"]["value"])
print(json.dumps(j["offline_c4"]))
PY
