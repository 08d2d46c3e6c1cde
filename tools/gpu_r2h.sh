#!/bin/bash
tag=${1:-r2h}
out=gpurun_out
mkdir -p $out
for sl in none 0.02 0.05 0.2 0.5; do
  if [ $sl = none ]; then unset CCB_SLACK; else export CCB_SLACK=$sl; fi
  timeout 300 python tools/tp_wall.py C2 1.0 --debuglib > $out/${tag}_tp_wall_c2_slack$sl.log 2>&1; echo "slack $sl"; tail -1 $out/${tag}_tp_wall_c2_slack$sl.log
done
for sl in none 0.05 0.2; do
  if [ $sl = none ]; then unset CCB_SLACK; else export CCB_SLACK=$sl; fi
  timeout 300 python tools/sweep.py --configs 4 --debuglib > $out/${tag}_sweep4_slack$sl.json 2>$out/${tag}_sweep4_slack$sl.err; echo "sweep slack $sl"
  python -c "
import json;j=json.load(open('$out/${tag}_sweep4_slack$sl.json'));print(j['seconds'],[(r['epsilon'],r['seconds'],r['blocks'],r['rounds']) for r in j['runs']])"
done
