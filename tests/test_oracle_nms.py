"""Pins oracle/nms_ref.{py,c} against the live third-party kernel (torchvision.ops.nms, present in the image)
and against the live reference `ops.non_max_suppression` (where /root/reference exists)."""
import numpy as np
import pytest
import torch
import torchvision

from oracle import nms_ref, refshim
from tests import nms_cases


def _tv(boxes, scores, thr):
    return torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()


@pytest.mark.parametrize("case", nms_cases.BOX_CASES, ids=lambda c: c[0])
def test_nms_matches_torchvision(case):
    name, fn, thr = case
    boxes, scores = fn()
    ref = _tv(boxes, scores, thr)
    got = nms_ref.nms(boxes, scores, thr)
    assert np.array_equal(ref, got), name
    if not np.isnan(scores).any():
        got_c = nms_ref.nms_c(boxes, scores, thr)
        assert np.array_equal(ref, got_c), name
        assert np.array_equal(ref[:7], nms_ref.nms_c(boxes, scores, thr, limit=7))


@pytest.mark.skipif(not refshim.available(), reason="live reference not present")
@pytest.mark.parametrize("case", nms_cases.PRED_CASES, ids=lambda c: c[0])
def test_nms_pipeline_matches_live_reference(case):
    name, fn, kw = case
    _, ops, _ = refshim.load()
    pred = fn()
    ref = ops.non_max_suppression(torch.from_numpy(pred.copy()), max_time_img=1e9, **kw)
    got = nms_ref.non_max_suppression(pred.copy(), **kw)
    assert len(ref) == len(got)
    for r, g in zip(ref, got):
        assert r.shape == g.shape, (name, r.shape, g.shape)
        assert np.array_equal(r.numpy(), g), name


def test_golden_nms_fixture():
    """Committed fixture generated from the live reference by tests/golden/make_golden.py."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "nms_golden.npz")
    g = np.load(path)
    for k in range(int(g["n_cases"])):
        name = str(g[f"name{k}"])
        case = {c[0]: c for c in nms_cases.PRED_CASES}[name]
        pred = case[1]()
        got = nms_ref.non_max_suppression(pred, **case[2])
        for b, det in enumerate(got):
            ref = g[f"c{k}_b{b}"]
            assert ref.shape == det.shape and np.array_equal(ref, det), (name, b)
