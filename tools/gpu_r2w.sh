#!/bin/bash
tag=${1:-r2w}
out=gpurun_out
mkdir -p $out
timeout 300 python tools/tp_wall.py C2 1.0 > $out/${tag}_tp_wall_c2.log 2>&1; tail -1 $out/${tag}_tp_wall_c2.log
timeout 300 python tools/tp_wall.py C2 1.0 49152 > $out/${tag}_tp_wall_c2_49152.log 2>&1; tail -1 $out/${tag}_tp_wall_c2_49152.log
timeout 300 python tools/tp_wall.py C2 1.0 65536 > $out/${tag}_tp_wall_c2_65536.log 2>&1; tail -1 $out/${tag}_tp_wall_c2_65536.log
timeout 300 python tools/tp_wall.py C2 1.0 24576 > $out/${tag}_tp_wall_c2_24576.log 2>&1; tail -1 $out/${tag}_tp_wall_c2_24576.log
timeout 300 python tools/tp_parts.py C2 1.0 > $out/${tag}_tp_parts_c2.log 2>&1; tail -1 $out/${tag}_tp_parts_c2.log
