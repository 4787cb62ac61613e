"""Golden outputs of the LIVE reference DetectionModel (imported via oracle/refshim.py) under the synthetic weights.
Small inputs only, so the fixtures stay small; committed because /root/reference does not travel to the GPU box."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import yolo_sod_b200  # noqa: E402,F401
from yolo_sod_b200 import cfg as ycfg, synth  # noqa: E402
from oracle import refshim  # noqa: E402

CASES = [  # (fixture key, config name, refshim name, batch, imgsz)
    ("sod64", "yolov12-sod-fusion-v5-simple", "sod", 1, 64),
    ("sod160", "yolov12-sod-fusion-v5-simple", "sod", 1, 160),
    ("v12n64", "yolov12n", "yolov12n", 1, 64),
    ("v12m64", "yolov12m", "yolov12m", 1, 64),
]


def main():
    out = {}
    for key, name, ref_name, b, sz in CASES:
        spec = ycfg.get_spec(name)
        sd = synth.synth_state_dict(spec, name, 0)
        m = refshim.build(ref_name)
        m.load_state_dict(sd, strict=True)
        x = synth.synth_images(b, sz, seed=7)
        feats = {}
        hooks = [mod.register_forward_hook(lambda mod_, i_, o_, idx=idx: feats.__setitem__(idx, o_))
                 for idx, mod in enumerate(m.model)]
        with torch.no_grad():
            y, raw = m(x)
        for h in hooks:
            h.remove()
        out[f"{key}_y"] = y.numpy().astype(np.float32)
        for l, r in enumerate(raw):
            out[f"{key}_raw{l}"] = r.numpy().astype(np.float16)   # fp16 storage: compared with tolerance anyway
        stats = []
        for idx in range(len(m.model) - 1):
            t = feats[idx]
            stats.append([float(t.mean()), float(t.std()), float(t.abs().max())])
        out[f"{key}_layer_stats"] = np.asarray(stats, np.float32)
    np.savez_compressed(os.path.join(HERE, "model_golden.npz"), **out)
    print("wrote model_golden.npz", os.path.getsize(os.path.join(HERE, "model_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
