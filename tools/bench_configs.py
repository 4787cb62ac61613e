"""Throughput / latency of the BASELINE.json configs C1..C5 on one B200 (device-resident inputs, CUDA events). Writes
gpurun_out/configs.json. (Parity of the 1024^2 config is a test: tests/test_gpu_model.py::test_sod_bf16_1024_single_image.)
    python tools/bench_configs.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_sod_b200  # noqa: E402,F401
from yolo_sod_b200 import ops, synth  # noqa: E402
from yolo_sod_b200.model import DetectionModel  # noqa: E402

SOD = "yolov12-sod-fusion-v5-simple"


def timed(fn, warm=3, n=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / n


def model_case(name, B, sz, n=10):
    m = DetectionModel(name, dtype=torch.bfloat16)
    xs = [synth.synth_images(B, sz, seed=i).cuda() for i in range(2)]
    k = [0]

    def step():
        k[0] += 1
        y, _ = m(xs[k[0] % 2])
        return ops.nms_padded(y, 0.25, 0.7, max_det=300)
    ms = timed(step, n=n)
    det, cnt, _ = step()
    del m
    torch.cuda.empty_cache()
    return {"model": name, "batch": B, "imgsz": sz, "ms_per_batch": round(ms, 4), "images_per_s": round(B / ms * 1e3, 1),
            "detections_per_image": round(float(cnt.float().mean()), 1)}


out = {}
out["C1_yolov12n_640_b1"] = model_case("yolov12n", 1, 640, n=50)
out["C2_sod_640_b32"] = model_case(SOD, 32, 640)
out["C3_sod_1024_b16"] = model_case(SOD, 16, 1024)
out["C4_sod_640_b256_one_gpu_8x32"] = model_case(SOD, 32, 640)   # literal C4 = the same model, 256 images sharded 8 x 32
out["C4b_yolov12m_640_b32"] = model_case("yolov12m", 32, 640)

# C5: NMS-only stress, 30k boxes x 10 classes, batch 64
g = torch.Generator().manual_seed(0)
B, A, nc = 64, 30000, 10
cxy = torch.rand(B, 2, A, generator=g) * 640
wh = torch.rand(B, 2, A, generator=g) * 60 + 2
cls = torch.zeros(B, nc, A)
idx = torch.randint(0, nc, (B, A), generator=g)
sc = torch.stack([torch.linspace(0.26, 0.999, A)[torch.randperm(A, generator=g)] for _ in range(B)])
cls.scatter_(1, idx[:, None, :], sc[:, None, :])
pred = torch.cat([cxy, wh, cls], 1).cuda()
ms = timed(lambda: ops.nms_padded(pred, 0.25, 0.7, max_det=300))
out["C5_nms_30k_x10_b64"] = {"ms_per_batch": round(ms, 4), "images_per_s": round(B / ms * 1e3, 1), "us_per_image": round(ms / B * 1e3, 2)}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
print(json.dumps(out, indent=1))
