"""GPU parity: csrc/nms.cu (through the C ABI via yolo_sod_b200.ops) vs the oracle -- bit-exact keep indices and rows."""
import os

import numpy as np
import pytest
import torch

import yolo_sod_b200  # noqa: F401
from oracle import nms_ref
from tests import nms_cases

pytestmark = pytest.mark.gpu


def _ops():
    from yolo_sod_b200 import ops
    return ops


@pytest.mark.parametrize("case", nms_cases.BOX_CASES, ids=lambda c: c[0])
def test_box_nms_bit_exact(case):
    name, fn, thr = case
    boxes, scores = fn()
    want = nms_ref.nms(boxes, scores, thr)
    got = _ops().nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), thr).cpu().numpy()
    assert np.array_equal(want, got), name
    if len(want) > 5:
        got5 = _ops().nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), thr, max_keep=5).cpu().numpy()
        assert np.array_equal(want[:5], got5)


@pytest.mark.parametrize("case", nms_cases.PRED_CASES + nms_cases.PRED_CASES_MULTILABEL, ids=lambda c: c[0])
def test_pipeline_bit_exact(case):
    name, fn, kw = case
    pred = fn()
    want, want_idx = nms_ref.non_max_suppression(pred.copy(), return_index=True, **kw)
    got = _ops().non_max_suppression(torch.from_numpy(pred).cuda(), **kw)
    assert len(want) == len(got)
    for w, g in zip(want, got):
        assert tuple(g.shape) == w.shape, (name, g.shape, w.shape)
        assert np.array_equal(g.cpu().numpy(), w), name
    kw2 = {k: v for k, v in kw.items()}
    det, count, index = _ops().nms_padded(torch.from_numpy(pred).cuda(), **kw2)
    for b, wi in enumerate(want_idx):
        n = int(count[b])
        assert n == len(wi)
        assert np.array_equal(index[b, :n].cpu().numpy().astype(np.int64), wi), name


def test_golden_fixture_from_live_reference():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "nms_golden.npz"))
    cases = nms_cases.PRED_CASES
    for k in range(int(g["n_cases"])):
        name, fn, kw = cases[k]
        assert str(g[f"name{k}"]) == name
        got = _ops().non_max_suppression(torch.from_numpy(fn()).cuda(), **kw)
        for b, d in enumerate(got):
            ref = g[f"c{k}_b{b}"]
            assert ref.shape == tuple(d.shape) and np.array_equal(ref, d.cpu().numpy()), (name, b)


def test_input_not_modified_and_asserts():
    pred = torch.from_numpy(nms_cases.make_pred(1, 10, 500, 3)).cuda()
    keep = pred.clone()
    _ops().non_max_suppression(pred, 0.25, 0.7)
    assert torch.equal(pred, keep)
    with pytest.raises(AssertionError):
        _ops().non_max_suppression(pred, 1.5, 0.7)
    with pytest.raises(AssertionError):
        _ops().non_max_suppression(pred, 0.25, -0.1)


@pytest.mark.parametrize("clustered", [False, True])
def test_stress_30k_vs_c_oracle(clustered):
    """BASELINE.json config 5 at reduced batch: 30 000 candidates x 10 classes, IoU 0.7, max_det 300."""
    pred = nms_cases.stress_pred(3, A=30000, clustered=clustered)
    want, want_idx = nms_ref.non_max_suppression(pred.copy(), 0.25, 0.7, max_det=300, return_index=True)
    det, count, index = _ops().nms_padded(torch.from_numpy(pred).cuda(), 0.25, 0.7, max_det=300)
    for b in range(3):
        n = int(count[b])
        assert n == len(want[b])
        assert np.array_equal(det[b, :n].cpu().numpy(), want[b])
        assert np.array_equal(index[b, :n].cpu().numpy().astype(np.int64), want_idx[b])


def test_c5_clustered_full_batch64_vs_c_oracle_on_8_images():
    """BASELINE config C5 at its stated size (B=64, 30 000 x 10, clustered = heavy suppression): the whole batch runs on the GPU;
    8 of the 64 images are checked bit-exactly (rows and keep indices) against the C oracle, all 64 by count sanity."""
    pred = nms_cases.stress_pred(64, A=30000, clustered=True)
    det, count, index = _ops().nms_padded(torch.from_numpy(pred).cuda(), 0.25, 0.7, max_det=300)
    torch.cuda.synchronize()
    cnt = count.cpu().numpy()
    assert (cnt > 0).all() and (cnt <= 300).all()
    for b in range(0, 64, 8):
        want, want_idx = nms_ref.non_max_suppression(pred[b:b + 1].copy(), 0.25, 0.7, max_det=300, return_index=True)
        n = int(cnt[b])
        assert n == len(want[0])
        assert np.array_equal(det[b, :n].cpu().numpy(), want[0])
        assert np.array_equal(index[b, :n].cpu().numpy().astype(np.int64), want_idx[0])


def test_zero_threshold_and_kept_list_size_boundaries():
    """(a) iou_thres == 0 must take the IEEE division for every pair -- a tiny intersection over a huge union underflows
    to IoU == 0, which is NOT > 0 (torchvision keeps the pair); (b) max_keep just below / at / above the kept-list kernel's
    shared-memory limit (1990) must all launch and agree with the oracle."""
    ops = _ops()
    big = np.float32(3.0e18)
    boxes = np.array([[0, 0, big, big], [0, 0, 1e-19, 1e-19], [5, 5, 6, 6]], np.float32)
    scores = np.array([0.9, 0.8, 0.7], np.float32)
    want = nms_ref.nms(boxes, scores, 0.0)
    got = ops.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), 0.0).cpu().numpy()
    assert np.array_equal(want, got), (want, got)
    b, s = nms_cases.uniform_boxes(6000, 77)
    for thr in (0.0, 1e-7, 0.3):
        want = nms_ref.nms(b, s, thr)
        got = ops.nms(torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda(), thr).cpu().numpy()
        assert np.array_equal(want, got), thr
    want = nms_ref.nms(b, s, 0.7)
    for mk in (1989, 1990, 1991, 2040, 2048, 2049):
        got = ops.nms(torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda(), 0.7, max_keep=mk).cpu().numpy()
        assert np.array_equal(want[:mk], got), mk


def test_full_size_properties_b64():
    """Config 5 at full size (B=64): size-independent properties instead of an oracle run."""
    pred = torch.from_numpy(nms_cases.stress_pred(64, A=30000, clustered=True)).cuda()
    det, count, index = _ops().nms_padded(pred, 0.25, 0.7, max_det=300)
    torch.cuda.synchronize()
    cnt = count.cpu().numpy()
    assert (cnt > 0).all() and (cnt <= 300).all()
    d = det.cpu().numpy()
    for b in range(0, 64, 7):
        n = cnt[b]
        s = d[b, :n, 4]
        assert (np.diff(s) <= 0).all(), "scores must be non-increasing"
        # idempotence: NMS over the kept boxes (with their class offsets) keeps every one of them
        boxes = d[b, :n, :4] + d[b, :n, 5:6] * np.float32(7680)
        again = nms_ref.nms_c(boxes, s, 0.7)
        assert np.array_equal(again, np.arange(n))
    # determinism
    det2, count2, _ = _ops().nms_padded(pred, 0.25, 0.7, max_det=300)
    assert torch.equal(det, det2) and torch.equal(count, count2)


def test_iou_threshold_boundary_is_bit_exact():
    """The greedy sweep classifies most pairs without the IEEE division (2^-20 guard band around the threshold). Thresholds
    placed exactly on, one ulp below and one ulp above the fp32 IoU of a pair must still agree with the oracle."""
    rng = np.random.default_rng(7)
    ops = _ops()
    bad = []
    for k in range(150):
        a = rng.uniform(0, 600, 2).astype(np.float32)
        wh = rng.uniform(5, 80, 2).astype(np.float32)
        d = (rng.uniform(-0.6, 0.6, 2) * wh).astype(np.float32)
        wh2 = (wh * rng.uniform(0.6, 1.5, 2)).astype(np.float32)
        boxes = np.stack([np.concatenate([a, a + wh]), np.concatenate([a + d, a + d + wh2])]).astype(np.float32)
        scores = np.array([0.9, 0.8], dtype=np.float32)
        x1, y1 = np.maximum(boxes[0, :2], boxes[1, :2])
        x2, y2 = np.minimum(boxes[0, 2:], boxes[1, 2:])
        w, h = np.float32(max(np.float32(0), np.float32(x2 - x1))), np.float32(max(np.float32(0), np.float32(y2 - y1)))
        inter = np.float32(w * h)
        area = lambda b: np.float32(np.float32(b[2] - b[0]) * np.float32(b[3] - b[1]))
        uni = np.float32(np.float32(area(boxes[0]) + area(boxes[1])) - inter)
        if not inter > 0:
            continue
        iou = np.float32(inter / uni)
        for thr in (float(iou), float(np.nextafter(iou, np.float32(0))), float(np.nextafter(iou, np.float32(1))),
                    float(iou) * (1 + 2.0 ** -21), float(iou) * (1 - 2.0 ** -21)):
            if not 0 <= thr <= 1:
                continue
            want = nms_ref.nms(boxes, scores, thr)
            got = ops.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), thr).cpu().numpy()
            if not np.array_equal(want, got):
                bad.append((k, thr, float(iou), want.tolist(), got.tolist()))
    assert not bad, bad[:5]
