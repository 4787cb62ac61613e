"""Timing of the CBAM statistics kernel with / without the cp.async prefetch ring (YSOD_RING_DEPTH = 0 | 8 | 16, read at first call).
    YSOD_RING_DEPTH=8 python tools/ab_ring.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_sod_b200  # noqa: E402,F401
from yolo_sod_b200 import lib  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (N, H, W, C) in [(32, 160, 160, 64), (32, 40, 40, 256), (16, 256, 256, 64)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).bfloat16()
    gate = torch.rand(N, C, device="cuda", generator=g)
    stats = torch.zeros(N, H * W, 2, device="cuda")
    ts = []
    for i in range(9):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lib.call("ysod_cbam_stats", lib.ptr(x), lib.BF16, N, H * W, C, C, lib.ptr(gate), lib.ptr(stats), lib.stream_ptr())
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    mb = x.numel() * 2 / 1e6
    t = sorted(ts)[len(ts) // 2]
    print(f"depth {os.environ.get('YSOD_RING_DEPTH', '0'):>2s}  {N}x{H}x{W}x{C}: {t:6.1f} us  {mb / t * 1e3 / 1e3:5.2f} TB/s  checksum {float(stats.double().sum()):.6f} {float(stats.abs().max()):.6f}")
