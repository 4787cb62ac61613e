"""Runs one tcgen05 conv shape a few times (for `ncu --set full -k regex:conv_tc_kernel`) and prints its event timing.

    python tools/prof_conv.py <N> <H> <W> <Cin> <Cout> <k> <s> [iters] [mode] [act]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_sod_b200  # noqa: E402,F401
from yolo_sod_b200 import lib  # noqa: E402

N, H, W, Cin, Cout, k, s = [int(v) for v in sys.argv[1:8]]
iters = int(sys.argv[8]) if len(sys.argv) > 8 else 5
mode = int(sys.argv[9]) if len(sys.argv) > 9 else 0
act = sys.argv[10] if len(sys.argv) > 10 else "silu"
pad = k // 2
Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
xcs, ocs = int(os.environ.get("PROF_XCS", Cin)), int(os.environ.get("PROF_OCS", Cout))   # pixel strides: channel slices of wider buffers
x = torch.randn(N, H, W, xcs, device="cuda").bfloat16()
cpad = (Cout + 15) // 16 * 16
w = (torch.randn(cpad, k * k * Cin, device="cuda") / (k * k * Cin) ** 0.5).bfloat16()
b = torch.zeros(cpad, device="cuda")
o = torch.empty(N, Ho, Wo, ocs, device="cuda", dtype=torch.bfloat16)
r = torch.randn(N, Ho, Wo, Cout, device="cuda").bfloat16() if os.environ.get("PROF_RES") else None
h = C.c_void_p()
lib.call("ysod_conv_tc_create_ex", C.byref(h), lib.ptr(x), N, H, W, Cin, xcs, lib.ptr(w), lib.ptr(b), Cout, cpad, k, s, lib.ptr(o), lib.BF16,
         ocs, lib.ptr(r) if r is not None else None, Cout if r is not None else 0, lib.ACT[act], mode)
info = (C.c_int * 8)()
lib.call("ysod_conv_tc_info", h, info)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(iters):
    if not os.environ.get("PROF_WARM"):
        flush.zero_()  # evict L2 between timed launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.call("ysod_conv_tc_run", h, lib.stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
fl = 2.0 * N * Ho * Wo * Cout * k * k * Cin
t = sorted(ts)[len(ts) // 2]
print(f"conv mode{mode} {Cin}->{Cout} k{k}s{s} N{N} {H}x{W}: tile {info[0]}x{info[1]} BN{info[2]} BK{info[3]} stages{info[4]} grid {info[5]}x{info[6]} "
      f"smem {info[7]}  median {t * 1e3:.1f} us  {fl / t / 1e9:.1f} TFLOP/s  ({fl / 1e9:.1f} GFLOP)")
