#!/bin/bash
# final 1-GPU pass of round 2: full GPU suite, smoke, both bench arms, app.run wall time, then the profiling pass
tag=${1:-r2z}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
tail -4 $out/${tag}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; echo "bench rc=$?"; tail -3 $out/${tag}_bench_c2.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err; echo "ref rc=$?"
timeout 600 python tools/app_wall.py 1.0 > $out/${tag}_app_wall_c2.json 2> $out/${tag}_app_wall_c2.err; echo "app_wall rc=$?"; cut -c1-900 $out/${tag}_app_wall_c2.json
timeout 300 python tools/tp_wall.py C2 1.0 > $out/${tag}_tp_wall_c2_graph.log 2>&1; tail -1 $out/${tag}_tp_wall_c2_graph.log
timeout 300 python tools/fp64_latency.py > $out/${tag}_fp64_latency.log 2>&1; cat $out/${tag}_fp64_latency.log
bash tools/gpu_profile_r2.sh $tag
timeout 300 python tools/trace_rounds.py C2 1.0 --tps 5 --out $out/${tag}_trace_c2.npz --detail 0 > $out/${tag}_trace_c2.log 2>&1; grep "== t" $out/${tag}_trace_c2.log
