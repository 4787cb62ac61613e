#!/bin/bash
# usage: tools/gpu_ab.sh "ENV=off" "ENV=on" [pytest -k expr]   -- model parity tests, then alternating quick C2 benches of the two settings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x ${3:+-k "$3"} 2>&1 | tail -n 4
for rep in 1 2 3; do
  for v in "$1" "$2"; do
    echo -n "$v: "
    env $v timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
  done
done
