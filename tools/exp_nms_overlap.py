"""Experiment: NMS of step i on a side stream while the forward of step i+1 runs (two model instances = two y buffers).
    python tools/exp_nms_overlap.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_sod_b200  # noqa: E402,F401
from yolo_sod_b200 import ops, synth  # noqa: E402
from yolo_sod_b200.model import DetectionModel  # noqa: E402

ms = [DetectionModel("yolov12-sod-fusion-v5-simple", dtype=torch.bfloat16) for _ in range(2)]
xs = [synth.synth_images(32, 640, seed=i).cuda() for i in range(4)]
side = torch.cuda.Stream()
done = [None, None]


def serial(i):
    y, _ = ms[0](xs[i % 4], static=True, want_raw=False)
    d, c, _ = ops.nms_padded(y, 0.25, 0.7, max_det=300)
    ops.clip_boxes(d, (640, 640))


def overlapped(i):
    k = i % 2
    main = torch.cuda.current_stream()
    if done[k] is not None:
        main.wait_event(done[k])
    y, _ = ms[k](xs[i % 4], static=True, want_raw=False)
    ev = torch.cuda.Event()
    ev.record(main)
    with torch.cuda.stream(side):
        side.wait_event(ev)
        d, c, _ = ops.nms_padded(y, 0.25, 0.7, max_det=300)
        ops.clip_boxes(d, (640, 640))
        done[k] = torch.cuda.Event()
        done[k].record(side)


def timed(fn, n=40):
    for i in range(6):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    torch.cuda.current_stream().wait_stream(side)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for rep in range(2):
    print("serial     : %.4f ms/step" % timed(serial))
    print("overlapped : %.4f ms/step" % timed(overlapped))
