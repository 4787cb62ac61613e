#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -n 4
for rep in 1 2; do
  for v in "a YSOD_NO_DUO=1" "b YSOD_NO_DUO=0" "c YSOD_C2F_CAT=0"; do
    set -- $v
    echo -n "$1 $2: "
    env $2 timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
  done
done
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/kernels_duo.json > /dev/null 2>&1
python - <<'PY'
import json
for f in ("kernels_duo",):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["sum_ms"])
    for o in d["per_op"]:
        if "L3:" in o["desc"] or "L27:" in o["desc"] or "L2:" in o["desc"]:
            print("  %.1f us  %s" % (o["ms"] * 1e3, o["desc"][:150]))
PY
