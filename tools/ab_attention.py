"""A/B of the attention core: mma.sync kernels (ysod_mha_core_ex impl 1) vs the tcgen05 / TMEM kernel (impl 2) on the shapes the
models launch at batch 32 (CUDA events, 20 runs after 5 warm-ups, inputs rotate over 4 buffers). Writes gpurun_out/ab_attention.json.
    python tools/ab_attention.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_sod_b200  # noqa: E402,F401
from yolo_sod_b200 import lib  # noqa: E402

SHAPES = [  # (what, batch, L, heads, D)
    ("SOD L9 SwinBlock P4 (36 windows x 32 img, 4 heads of 64)", 36 * 32, 49, 4, 64),
    ("SOD L12 A2_Attn (8 x 20 tokens, 8 heads of 64)", 32, 160, 8, 64),
    ("SOD-1024 L9 SwinBlock P4 (100 windows x 16 img)", 100 * 16, 49, 4, 64),
    ("yolov12n L6 AAttn 40x40 area 4 (2 heads of 32)", 32 * 4, 400, 2, 32),
    ("yolov12n L8 AAttn 20x20 area 1 (4 heads of 32)", 32, 400, 4, 32),
    ("yolov12m L6 AAttn 40x40 area 4 (8 heads of 32)", 32 * 4, 400, 8, 32),
    ("unfused SOD L28 SwinBlock P2 (529 windows x 32 img, 2 heads of 32)", 529 * 32, 49, 2, 32),
]
out = []
for what, batch, L, heads, D in SHAPES:
    E = heads * D
    bufs = [torch.randn(batch, L, 3 * E, device="cuda").bfloat16() for _ in range(4)]
    o = torch.empty(batch, L, E, dtype=torch.bfloat16, device="cuda")
    row = {"shape": what, "batch": batch, "L": L, "heads": heads, "D": D, "gflop": 4.0 * batch * heads * L * L * D / 1e9}
    res = {}
    for impl, name in ((1, "mma_sync"), (2, "tcgen05")):
        def run(i):
            d = bufs[i % 4]
            lib.call("ysod_mha_core_ex", lib.ptr(d), lib.ptr(d, E), lib.ptr(d, 2 * E), lib.BF16, batch, L, heads, D, 3 * E, 3 * E, 3 * E,
                     L * 3 * E, L * 3 * E, L * 3 * E, 1.0 / D ** 0.5, lib.ptr(o), E, L * E, impl, lib.stream_ptr())
        for i in range(5):
            run(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(20):
            run(i)
        b.record()
        b.synchronize()
        row[name + "_us"] = round(a.elapsed_time(b) / 20 * 1e3, 2)
        res[name] = o.float().clone()
    row["max_abs_diff"] = float((res["mma_sync"] - res["tcgen05"]).abs().max())
    row["speedup_tcgen05"] = round(row["mma_sync_us"] / row["tcgen05_us"], 3)
    out.append(row)
    print(row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ab_attention.json", "w"), indent=1)
