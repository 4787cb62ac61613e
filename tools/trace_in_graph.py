"""Pipeline timeline of CTA 0 of ONE tcgen05 conv op while it runs inside the captured forward graph (warm L2, PDL-chained neighbours):
    YSOD_TRACE_OP=<op index> python tools/trace_in_graph.py [batch] [imgsz]
Prints kernel entry -> prologue done -> first operands landed -> ... -> all roles done in SM cycles (see tools/trace_conv.py)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_sod_b200  # noqa: E402,F401
from yolo_sod_b200 import lib, synth  # noqa: E402
from yolo_sod_b200.model import DetectionModel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sz = int(sys.argv[2]) if len(sys.argv) > 2 else 640
op = int(os.environ["YSOD_TRACE_OP"])
m = DetectionModel("yolov12-sod-fusion-v5-simple", dtype=torch.bfloat16)
x = synth.synth_images(B, sz, seed=1).cuda()
prog = m.program(B, sz, sz, False, False)
print("op", op, prog.op_desc[op])
buf = np.zeros(2 * 8192, dtype=np.uint64)
for it in range(3):
    m(x, static=True, want_raw=False)
    n = lib.load().ysod_debug_trace(buf.ctypes.data_as(C.c_void_p), 8192)
rec = [(int(buf[2 * i + 1]), int(buf[2 * i]) >> 56, (int(buf[2 * i]) >> 48) & 0xff, (int(buf[2 * i]) >> 32) & 0xffff, int(buf[2 * i]) & 0xffffffff)
       for i in range(n) if buf[2 * i + 1] != 0]
rec.sort()
t0 = rec[0][0]
names = {(3, 1): "M1 acc free", (3, 2): "M1 operands landed", (3, 3): "M1 burst issued", (0, 1): "P slot free", (1, 1): "M acc free", (1, 2): "M operands landed",
         (1, 3): "M burst issued", (2, 1): "E acc full", (2, 10): "CTA kernel entry", (2, 11): "CTA prologue done", (2, 12): "CTA all roles done",
         (2, 2): "E barrier1", (2, 3): "E acc released", (2, 4): "E barrier2"}
first = {}
last = {}
for t, role, ev, tile, idx in rec:
    first.setdefault((role, ev), t - t0)
    last[(role, ev)] = t - t0
for k in sorted(first):
    print(f"  {names.get(k, k):24s} first {first[k]:8d}  last {last[k]:8d}")
