"""Generates tests/golden/tiny_ckpt.pt: a reference-format checkpoint (pickled `ultralytics.nn.tasks.DetectionModel`, fp16, the dict
layout of engine/trainer.py save_model) of a tiny SOD-style model, written by the LIVE reference in the build container.
The checkpoint loader under test (yolo_sod_b200/checkpoint.py) must read it WITHOUT the reference installed.

    python tests/golden/make_golden_ckpt.py
"""
import io
import os
import sys
from copy import deepcopy

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

TINY = {
    "nc": 3, "depth_multiple": 1.0, "width_multiple": 0.25, "ch": 3,
    "backbone": [
        [-1, 1, "Conv", [64, 3, 2]], [-1, 1, "SE_Block", [16]], [-1, 1, "Conv", [64, 3, 2]], [-1, 1, "C2f", [64, True]],
        [-1, 1, "CBAM_Block", [64, 8]], [-1, 1, "Conv", [128, 3, 2]], [-1, 1, "C2f", [128, True]],
    ],
    "neck": [
        [-1, 1, "Conv", [64, 1, 1]], [-1, 1, "nn.Upsample", [None, 2, "nearest"]], [[-1, 4], 1, "Concat", [1]], [-1, 1, "C2f", [64, False]],
        [-1, 1, "CA_Block", [64]],
    ],
    "head": [[[11, 6], 1, "Detect", ["nc"]]],
}
OUT = os.path.join(ROOT, "tests", "golden", "tiny_ckpt.pt")


def main():
    DetectionModel, _, _ = refshim.load()
    torch.manual_seed(0)
    m = DetectionModel(deepcopy(TINY), verbose=False)
    g = torch.Generator().manual_seed(1)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data.uniform_(0.75, 1.25, generator=g)
            mod.bias.data.normal_(0, 0.2, generator=g)
            mod.running_mean.normal_(0, 0.1, generator=g)
            mod.running_var.uniform_(0.2, 0.6, generator=g)
    m.names = {0: "pedestrian", 1: "car", 2: "bicycle"}
    m.eval()
    ckpt = {"epoch": 3, "best_fitness": None, "model": None, "ema": deepcopy(m).half(), "updates": 12, "optimizer": None,
            "train_args": {"imgsz": 640, "batch": 4, "model": "tiny-sod.yaml"}, "date": "2026-10-18T00:00:00", "version": "8.3.63"}
    buf = io.BytesIO()
    torch.save(ckpt, buf)
    open(OUT, "wb").write(buf.getvalue())
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", sum(p.numel() for p in m.parameters()), "parameters;", len(m.state_dict()), "tensors")


if __name__ == "__main__":
    main()
