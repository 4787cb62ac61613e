#!/bin/bash
# resident-weight plan of the generic conv kernel: unit tests, then in-graph traces of deep 1x1 layers with the plan on / off
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu -x 2>&1 | tail -n 4
python - <<'PY'
import torch, yolo_sod_b200
from yolo_sod_b200.model import DetectionModel
m = DetectionModel("yolov12-sod-fusion-v5-simple", dtype=torch.bfloat16)
prog = m.program(32, 640, 640, False, False)
for i, d in enumerate(prog.op_desc):
    if " tc " in d and ("k1s1" in d or "k3s2" in d): print(i, d[:140])
PY
for op in ${OPS:-52 22}; do
  for g in 1 0; do
    echo "== op $op YSOD_NO_GRES=$g"
    YSOD_NO_GRES=$g YSOD_TRACE_OP=$op timeout 120 python tools/trace_in_graph.py 2>&1 | tail -n 14 | cut -c1-200
  done
done
