#!/bin/bash
tag=${1:-r2i}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
tail -4 $out/${tag}_pytest_gpu.log
timeout 300 python tools/bench_offline.py --M 100000 --D 40 > $out/${tag}_offline_c4_1gpu.json 2>&1; cat $out/${tag}_offline_c4_1gpu.json | cut -c1-900
