"""Experiment: one 32-image forward graph vs two concurrent 16-image graphs on two streams (do interleaved chains hide the
per-kernel ramp / drain?).  python tools/exp_two_chains.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_sod_b200  # noqa: E402,F401
from yolo_sod_b200 import synth  # noqa: E402
from yolo_sod_b200.model import DetectionModel  # noqa: E402

m = DetectionModel("yolov12-sod-fusion-v5-simple", dtype=torch.bfloat16)
m2 = DetectionModel("yolov12-sod-fusion-v5-simple", dtype=torch.bfloat16)
x = synth.synth_images(32, 640, seed=1).cuda()
xa, xb = x[:16].contiguous(), x[16:].contiguous()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def one():
    m(x, static=True, want_raw=False)


def two():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        m(xa, static=True, want_raw=False)
    with torch.cuda.stream(s2):
        m2(xb, static=True, want_raw=False)
    cur.wait_stream(s1); cur.wait_stream(s2)


def seq():
    m(xa, static=True, want_raw=False)
    m2(xb, static=True, want_raw=False)


print("one graph B=32           : %.3f ms" % timed(one))
print("two graphs B=16 sequential: %.3f ms" % timed(seq))
print("two graphs B=16 concurrent: %.3f ms" % timed(two))
