"""A/B of the memory-bound block kernels on the SOD-simple shapes at batch 32 (CUDA events, L2 flushed between runs):
CBAM spatial attention as stats + apply passes vs the single-pass kernel; CoordAtt pooling two-pass vs single-pass.
    python tools/ab_blocks.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_sod_b200  # noqa: E402,F401
from yolo_sod_b200 import lib  # noqa: E402

L = lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, n=7):
    ts = []
    for i in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]


out = []
st = lib.stream_ptr
for (N, H, W, C) in [(32, 160, 160, 64), (32, 40, 40, 256), (16, 256, 256, 64)]:
    x = torch.randn(N, H, W, C, device="cuda").bfloat16()
    o1, o2 = torch.empty_like(x), torch.empty_like(x)
    gate = torch.rand(N, C, device="cuda")
    stats = torch.empty(N, H * W, 2, device="cuda")
    wsp = torch.randn(2, 7, 7, device="cuda") * 0.2

    def two_pass():
        lib.call("ysod_cbam_stats", lib.ptr(x), lib.BF16, N, H * W, C, C, lib.ptr(gate), lib.ptr(stats), st())
        lib.call("ysod_cbam_apply", lib.ptr(x), lib.BF16, N, H, W, C, C, lib.ptr(gate), lib.ptr(stats), lib.ptr(wsp), 7, lib.ptr(o1), C, st())

    def one_pass():
        lib.call("ysod_cbam_spatial", lib.ptr(x), lib.BF16, N, H, W, C, C, lib.ptr(gate), lib.ptr(wsp), 7, lib.ptr(o2), C, st())
    t2, t1 = timed(two_pass), timed(one_pass)
    mb = 2 * x.numel() * 2 / 1e6
    row = {"kernel": "CBAM spatial", "shape": [N, H, W, C], "two_pass_us": round(t2, 1), "one_pass_us": round(t1, 1), "equal": bool(torch.equal(o1, o2)),
           "map_rw_MB": mb, "one_pass_GBs": round(mb / t1 * 1e3, 0)}
    out.append(row)
    print(row, flush=True)
for (N, H, W, C) in [(32, 80, 80, 128), (16, 128, 128, 128)]:
    x = torch.randn(N, H, W, C, device="cuda").bfloat16()
    p1, p2 = torch.empty(N, H + W, C, device="cuda"), torch.empty(N, H + W, C, device="cuda")
    nws = int(L.ysod_ca_pool_workspace_floats(N, H, W, C))
    ws = torch.empty(max(nws, 1), device="cuda")
    t2 = timed(lambda: lib.call("ysod_ca_pool", lib.ptr(x), lib.BF16, N, H, W, C, C, lib.ptr(p1), None, st()))
    t1 = timed(lambda: lib.call("ysod_ca_pool", lib.ptr(x), lib.BF16, N, H, W, C, C, lib.ptr(p2), lib.ptr(ws), st()))
    row = {"kernel": "CoordAtt pool", "shape": [N, H, W, C], "two_pass_us": round(t2, 1), "one_pass_us": round(t1, 1),
           "max_abs_diff": float((p1 - p2).abs().max()), "map_MB": x.numel() * 2 / 1e6}
    out.append(row)
    print(row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ab_blocks.json", "w"), indent=1)
