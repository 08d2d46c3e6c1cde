#!/bin/bash
tag=${1:-r2e}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
tail -4 $out/${tag}_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; echo "bench rc=$?"; tail -3 $out/${tag}_bench_c2.err; cut -c1-3000 $out/${tag}_bench_c2.json
timeout 600 python tools/app_wall.py 1.0 > $out/${tag}_app_wall_c2.json 2> $out/${tag}_app_wall_c2.err; echo "app_wall rc=$?"; tail -3 $out/${tag}_app_wall_c2.err; cat $out/${tag}_app_wall_c2.json | cut -c1-1500
