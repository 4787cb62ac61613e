#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -n 5
for rep in 1 2; do
  for v in "a YSOD_SWIN_IMPL=1" "b YSOD_SWIN_IMPL=0"; do
    set -- $v
    echo -n "$1 $2: "
    env $2 timeout 300 python bench.py --quick --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
  done
done
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/kernels_swintc.json > /dev/null 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/kernels_swintc.json"))
print(d["sum_ms"])
for o in d["per_op"]:
    if "L28:" in o["desc"] or "L9:" in o["desc"]:
        print("  %.1f us  %s" % (o["ms"] * 1e3, o["desc"][:150]))
PY
