"""Event timing of ysod_dwconv3_ln (SwinBlock dw 3x3 + LayerNorm 1 with the tokens in NHWC order) on an N x H x W x C map.
    python tools/prof_dwln.py [N] [H] [W] [C] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_sod_b200  # noqa: E402,F401
from yolo_sod_b200 import lib  # noqa: E402

N, H, W, C = [int(v) for v in sys.argv[1:5]] if len(sys.argv) > 4 else (32, 40, 40, 256)
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 9
g = torch.Generator().manual_seed(5)
x = torch.randn(N, H, W, C, generator=g).bfloat16().cuda()
w = (torch.randn(9, C, generator=g) / 3).cuda()
gamma, beta = (1 + 0.1 * torch.randn(C, generator=g)).cuda(), (0.1 * torch.randn(C, generator=g)).cuda()
y, yn = torch.empty_like(x), torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.call("ysod_dwconv3_ln", lib.ptr(x), N, H, W, C, C, lib.ptr(w), lib.ptr(gamma), lib.ptr(beta), 1e-5, lib.ptr(y), C, lib.ptr(yn), C, lib.stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = sorted(ts)[len(ts) // 2]
print(f"dwconv3_ln N{N} {H}x{W}x{C}: median {t * 1e3:.1f} us  {3 * x.numel() * 2 / t / 1e6:.0f} GB/s  checksums {float(y.float().sum()):.2f} {float(yn.float().abs().sum()):.1f}")
