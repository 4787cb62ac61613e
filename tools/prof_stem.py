"""Event timing of the tensor-core stem (ysod_stem_mma / ysod_stem_mma_gap) on a B x 3 x S x S batch, fp32 and uint8 sources.
    python tools/prof_stem.py [B] [S] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_sod_b200  # noqa: E402,F401
from yolo_sod_b200 import lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 640
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 9
g = torch.Generator().manual_seed(3)
xf = torch.rand(B, 3, S, S, generator=g).cuda()
xu = (torch.rand(B, S, S, 3, generator=g) * 255).to(torch.uint8).cuda()
wk = torch.zeros(32, 32)
wk[:, :27] = torch.randn(32, 27, generator=g) / 5
wk = wk.bfloat16().cuda()
bias = (torch.randn(32, generator=g) * 0.1).cuda()
out = torch.empty(B, S // 2, S // 2, 32, dtype=torch.bfloat16, device="cuda")
Sx = -(-(S // 2) // 64) * -(-(S // 2) // 4)
psum = torch.empty(B, Sx, 32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, src, fmt in (("fp32 NCHW", xf, 0), ("uint8 BGR HWC", xu, 1)):
    for gap in (0, 1):
        ts = []
        for i in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if gap:
                lib.call("ysod_stem_mma_gap", lib.ptr(src), fmt, B, S, S, lib.ptr(wk), lib.ptr(bias), 32, lib.ptr(out), 32, lib.ACT["silu"], lib.ptr(psum), lib.stream_ptr())
            else:
                lib.call("ysod_stem_mma", lib.ptr(src), fmt, B, S, S, lib.ptr(wk), lib.ptr(bias), 32, lib.ptr(out), 32, lib.ACT["silu"], lib.stream_ptr())
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        nbytes = src.numel() * src.element_size() + out.numel() * 2
        print(f"stem {name}{' + GAP partials' if gap else ''}: B{B} {S}x{S} median {t * 1e3:.1f} us  {nbytes / t / 1e6:.0f} GB/s  checksum {float(out.float().sum()):.1f}")
