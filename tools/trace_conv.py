"""Pipeline timeline of CTA 0 of the tcgen05 conv kernel (debug bit 32).  python tools/trace_conv.py <N> <H> <W> <Cin> <Cout> <k> <s> [extra dbg bits]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_sod_b200  # noqa: E402,F401
from yolo_sod_b200 import lib  # noqa: E402

N, H, W, Cin, Cout, k, s = [int(v) for v in sys.argv[1:8]]
extra = int(sys.argv[8]) if len(sys.argv) > 8 else 0
pad = k // 2
Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
x = torch.randn(N, H, W, Cin, device="cuda").bfloat16()
cpad = (Cout + 15) // 16 * 16
w = (torch.randn(cpad, k * k * Cin, device="cuda") / (k * k * Cin) ** 0.5).bfloat16()
b = torch.zeros(cpad, device="cuda")
o = torch.empty(N, Ho, Wo, Cout, device="cuda", dtype=torch.bfloat16)
h = C.c_void_p()
lib.call("ysod_conv_tc_create_ex", C.byref(h), lib.ptr(x), N, H, W, Cin, Cin, lib.ptr(w), lib.ptr(b), Cout, cpad, k, s, lib.ptr(o), lib.BF16,
         Cout, None, 0, lib.ACT["silu"], (32 | extra) << 8)
buf = np.zeros(2 * 8192, dtype=np.uint64)
for it in range(2):
    lib.call("ysod_conv_tc_run", h, lib.stream_ptr())
    n = lib.load().ysod_debug_trace(buf.ctypes.data_as(C.c_void_p), 8192)
rec = [(int(buf[2 * i + 1]), int(buf[2 * i]) >> 56, (int(buf[2 * i]) >> 48) & 0xff, (int(buf[2 * i]) >> 32) & 0xffff, int(buf[2 * i]) & 0xffffffff)
       for i in range(n) if buf[2 * i + 1] != 0]
rec.sort()
t0 = rec[0][0]
names = {(3, 1): "M1 acc free", (3, 2): "M1 operands landed", (3, 3): "M1 burst issued", (0, 1): "P  slot free", (1, 1): "M  acc free", (1, 2): "M  operands landed", (1, 3): "M  burst issued", (2, 1): "E  acc full",
         (1, 5): "M  mmas issued", (1, 6): "M  reconverged", (1, 7): "M  loop top", (2, 10): "CTA kernel entry", (2, 11): "CTA prologue done", (2, 12): "CTA all roles done", (2, 2): "E  barrier1", (2, 3): "E  acc released", (2, 4): "E  barrier2"}
print(f"{n} records; extra dbg {extra}")
for t, role, ev, tile, idx in rec:
    if (0 if len(sys.argv) > 9 else 3) <= tile <= 7 or ev >= 10:
        col = {0: 0, 1: 1, 3: 2, 2: 3}[role]
        print(f"{t - t0:8d}  {' ' * 22 * col}{names.get((role, ev), (role, ev))} t{tile} #{idx}")

# steady-state summary: per role, time between consecutive tiles at the role's last event
import collections
last = collections.defaultdict(dict)
for t, role, ev, tile, idx in rec:
    last[(role, ev)][tile] = t
for (role, ev), d in sorted(last.items()):
    tiles = sorted(d)
    if len(tiles) < 4:
        continue
    ts = [d[k] for k in tiles]
    deltas = [b - a for a, b in zip(ts, ts[1:])]
    print(f"role {role} ev {ev}: {len(tiles)} tiles, first at {ts[0] - t0}, last at {ts[-1] - t0}, per-own-tile delta min {min(deltas)} "
          f"median {sorted(deltas)[len(deltas) // 2]} max {max(deltas)}; deltas[:12]={deltas[:12]} deltas[-6:]={deltas[-6:]}")
