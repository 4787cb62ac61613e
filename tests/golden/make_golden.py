"""Generates tests/golden/*.npz by running the LIVE reference (/root/reference, via oracle/refshim.py) in the
build container. The reference cannot travel to the GPU box, so its outputs are committed as small fixtures.

    python tests/golden/make_golden.py [nms] [model]
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from tests import nms_cases  # noqa: E402


def make_nms():
    _, ops, _ = refshim.load()
    out = {}
    cases = nms_cases.PRED_CASES + nms_cases.PRED_CASES_MULTILABEL
    out["n_cases"] = np.int64(len(nms_cases.PRED_CASES))
    for k, (name, fn, kw) in enumerate(cases):
        pred = fn()
        dets = ops.non_max_suppression(torch.from_numpy(pred.copy()), max_time_img=1e9, **kw)
        key = f"c{k}" if k < len(nms_cases.PRED_CASES) else f"ml{k - len(nms_cases.PRED_CASES)}"
        out[f"name{k}"] = np.array(name)
        for b, d in enumerate(dets):
            out[f"{key}_b{b}"] = d.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "nms_golden.npz"), **out)
    print("wrote nms_golden.npz", sum(v.nbytes for v in out.values()), "bytes")


if __name__ == "__main__":
    what = sys.argv[1:] or ["nms", "model"]
    if "nms" in what:
        make_nms()
    if "model" in what:
        from tests.golden import make_golden_model
        make_golden_model.main()
