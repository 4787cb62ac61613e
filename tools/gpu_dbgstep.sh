#!/bin/bash
# upper bounds inside the captured forward: conv epilogue without work (1) / without the TMA store (8) / producer without TMA loads (2) / no MMA (4)
for dbg in 0 8 1 2 4 0; do
  echo -n "YSOD_CONV_DEBUG=$dbg: "
  YSOD_CONV_DEBUG=$dbg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('step_split_ms'))"
done
