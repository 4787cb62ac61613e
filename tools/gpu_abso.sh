#!/bin/bash
# same-box A/B of differently built libraries: tools/gpu_abso.sh libysod_varA.so libysod.so ...   (paths relative to yolo-sod_b200/)
for rep in 1 2 3; do
  for so in "$@"; do
    echo -n "$so: "
    YSOD_LIB_PATH=$PWD/yolo-sod_b200/$so timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
  done
done
