"""Runs the fused P2 SwinBlock kernels (tcgen05: ysod_swin64_tc, mma.sync: ysod_swin64_fused) on an N x H x W x 64 map and prints their
event timings (for `ncu -k regex:swin64`).   python tools/prof_swin.py [N] [H] [W] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_sod_b200  # noqa: E402,F401
from yolo_sod_b200 import lib  # noqa: E402

N, H, W = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (32, 160, 160)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
g = torch.Generator().manual_seed(7)
x = torch.randn(N, H, W, 64, generator=g).bfloat16().cuda()
wb = torch.cat([torch.randn(576, generator=g) / 3, torch.randn(192 * 64, generator=g) / 8, torch.randn(64 * 64, generator=g) / 8,
                torch.randn(128 * 64, generator=g) / 8, torch.randn(64 * 128, generator=g) / 11, torch.randn(64 * 64, generator=g) / 8]).bfloat16().cuda()
pf = torch.randn(768, generator=g) * 0.1
pf[192:320] = 0.0
pf = pf.cuda()
o = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name in ("ysod_swin64_fused", "ysod_swin64_tc"):
    ts = []
    for i in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.call(name, lib.ptr(x), N, H, W, 64, lib.ptr(wb), lib.ptr(pf), lib.ptr(o), 64, 7, 2, lib.stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name}: N{N} {H}x{W} median {sorted(ts)[len(ts) // 2] * 1e3:.1f} us")

# stage trace of the tcgen05 kernel (CTA 0, group 0, thread 0; SM cycles)
import ctypes
LABELS = ["A patches->smem", "B dw 3x3", "B LN1", "QKV mma wait", "C qkv epilogue", "S mma wait", "D softmax+PV (both heads)", "E O->A1", "out_proj wait",
          "F +res LN2", "MLP1 wait", "G GELU", "MLP2 wait (+identity load)", "H +res", "pw wait", "I SiLU+store"]
buf = (ctypes.c_longlong * (8 * 24))()
lib.call("ysod_swin64_tc_trace", 1, None)
lib.call("ysod_swin64_tc", lib.ptr(x), N, H, W, 64, lib.ptr(wb), lib.ptr(pf), lib.ptr(o), 64, 7, 2, lib.stream_ptr())
torch.cuda.synchronize()
lib.call("ysod_swin64_tc_trace", 0, ctypes.cast(buf, ctypes.c_void_p))
v = list(buf)
for tile in range(2, 6):
    st = v[tile * 24: tile * 24 + 17]
    print(f"tile {tile}: total {st[16] - st[0]} cycles; next tile starts +{v[(tile + 1) * 24] - st[16]}")
    print("   " + " | ".join(f"{LABELS[i]} {st[i + 1] - st[i]}" for i in range(16)))
