"""Produces yolo-sod_b200/data/synth_calib.json: per-BN-layer scalar (mean, var) of the conv output under the synthetic
weights, and the per-level class-bias shift. Uses the oracle forward (test infrastructure) with its BN hook patched so that
every BN is calibrated on inputs produced by already-calibrated earlier layers (one sequential pass).

    python tests/golden/make_synth_calib.py   (fixture generator: lives with the other golden generators because it imports oracle/)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import yolo_sod_b200  # noqa: E402,F401
from yolo_sod_b200 import cfg as ycfg, synth  # noqa: E402
from oracle import model_ref  # noqa: E402

CONFIGS = ["yolov12-sod-fusion-v5-simple", "yolov12-sod-fusion-v5-stable", "yolov12-sod-fusion-v5", "yolov12n", "yolov12s", "yolov12m", "E1", "E2", "E3", "E4", "E5", "E6"]


def calibrate(name, seed=0, target_frac=0.04):
    spec = ycfg.get_spec(name)
    strides = ycfg.strides_of(spec)
    sd = synth.synth_state_dict(spec, name, seed, calib={})
    table = {}
    orig_bn = model_ref._bn

    def bn_hook(sd_, pfx, x):
        m, v = float(x.mean()), float(x.var())
        table[pfx] = [round(m, 6), round(max(v, 1e-6), 6)]
        sd_[f"{pfx}.running_mean"].fill_(table[pfx][0])
        sd_[f"{pfx}.running_var"].fill_(table[pfx][1])
        return orig_bn(sd_, pfx, x)

    model_ref._bn = bn_hook
    try:
        x = synth.synth_images(2, 640, seed=99)
        # zero class bias for the quantile measurement
        for k in sd:
            if ".cv3." in k and k.endswith(".2.bias"):
                sd[k].zero_()
        y, raw = model_ref.forward(spec, sd, x, strides)
    finally:
        model_ref._bn = orig_bn
    nc = spec.nc
    cls_bias = []
    for r in raw:
        mx = r[:, 64:64 + nc].amax(1).flatten()
        q = torch.quantile(mx, 1.0 - target_frac)
        cls_bias.append(round(float(-1.0986 - q), 4))   # sigmoid(x) > 0.25  <=>  x > -1.0986
    table["cls_bias"] = cls_bias
    return table


if __name__ == "__main__":
    path = os.path.join(ROOT, "yolo-sod_b200", "data", "synth_calib.json")
    only = sys.argv[1:]                      # optional: recompute just these configs, keep the rest of the table
    out = json.load(open(path)) if (only and os.path.exists(path)) else {}
    for name in (only or CONFIGS):
        out[synth.calib_key(name, 0)] = calibrate(name, 0)
        print(name, "BN layers:", len(out[synth.calib_key(name, 0)]) - 1, "cls_bias:", out[synth.calib_key(name, 0)]["cls_bias"])
    json.dump(out, open(path, "w"), indent=0, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")
