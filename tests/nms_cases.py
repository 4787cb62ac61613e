"""Seeded NMS inputs shared by the oracle tests, the GPU parity tests and the golden generator.

Everything is numpy RandomState-seeded so the same arrays are rebuilt on the GPU box without /root/reference.
Distributions follow SURVEY.md section 8(d): uniform boxes (little overlap) and jittered clusters (heavy suppression).
"""
import numpy as np

F32 = np.float32


def uniform_boxes(n, seed, size=640.0, wh=(2.0, 62.0), distinct_scores=True, lo=0.26):
    r = np.random.RandomState(seed)
    cxy = r.uniform(0, size, (n, 2))
    wh_ = r.uniform(wh[0], wh[1], (n, 2))
    b = np.concatenate([cxy - wh_ / 2, cxy + wh_ / 2], 1).astype(F32)
    if distinct_scores:
        s = np.linspace(lo, 0.999, n).astype(F32)[r.permutation(n)]
    else:
        s = r.uniform(lo, 1, n).astype(F32)
    return b, s


def clustered_boxes(n, seed, centres=60, size=640.0, jitter=4.0):
    r = np.random.RandomState(seed)
    c = r.uniform(20, size - 20, (centres, 2))
    whc = r.uniform(8, 60, (centres, 2))
    k = r.randint(0, centres, n)
    cxy = c[k] + r.uniform(-jitter, jitter, (n, 2))
    wh = whc[k] * r.uniform(0.85, 1.15, (n, 2))
    b = np.concatenate([cxy - wh / 2, cxy + wh / 2], 1).astype(F32)
    s = np.linspace(0.26, 0.999, n).astype(F32)[r.permutation(n)]
    return b, s


def _ties():
    b, s = clustered_boxes(1500, 3)
    s = (np.round(s * 20) / 20).astype(F32)  # many equal scores
    return b, s


def _exact_half():
    b = np.array([[0, 0, 2, 1], [0, 0, 1, 1], [10, 10, 12, 12], [10, 10, 12, 11]], F32)
    s = np.array([0.9, 0.8, 0.7, 0.6], F32)
    return b, s


def _exact_point6():
    # IoU = 3/5 -> fp32 0.6f, which as double is > 0.6: suppressed by the reference's double compare
    b = np.array([[0, 0, 5, 1], [0, 0, 3, 1], [0, 0, 5, 1]], F32)
    s = np.array([0.9, 0.8, 0.7], F32)
    return b, s


def _degenerate():
    b = np.array([[5, 5, 5, 5], [5, 5, 5, 5], [0, 0, 0, 10], [0, 0, 0, 10], [3, 3, 1, 1], [3, 3, 1, 1],
                  [0, 0, 10, 10], [1, 1, 9, 9]], F32)
    s = np.array([0.9, 0.8, 0.7, 0.6, 0.5, 0.45, 0.4, 0.3], F32)
    return b, s


def _nan_coord():
    b, s = uniform_boxes(200, 5)
    b[17, 2] = np.nan
    b[40] = np.nan
    return b, s


def _big_offset():
    b, s = clustered_boxes(1200, 7)
    r = np.random.RandomState(8)
    c = r.randint(0, 80, (1200, 1)).astype(F32) * F32(7680)
    return (b + c).astype(F32), s


def _one():
    return np.array([[1, 2, 3, 4]], F32), np.array([0.5], F32)


def _empty():
    return np.zeros((0, 4), F32), np.zeros((0,), F32)


BOX_CASES = [
    ("uniform2k_iou0.7", lambda: uniform_boxes(2000, 0), 0.7),
    ("uniform2k_dupscores_iou0.45", lambda: uniform_boxes(2000, 1, distinct_scores=False), 0.45),
    ("clustered3k_iou0.7", lambda: clustered_boxes(3000, 2), 0.7),
    ("clustered3k_iou0.3", lambda: clustered_boxes(3000, 2), 0.3),
    ("ties", _ties, 0.5),
    ("iou_equals_thr_kept", _exact_half, 0.5),
    ("iou_point6_double_compare", _exact_point6, 0.6),
    ("degenerate_zero_area_nan_iou", _degenerate, 0.5),
    ("nan_coordinates", _nan_coord, 0.5),
    ("class_offset_precision", _big_offset, 0.7),
    ("single", _one, 0.5),
    ("empty", _empty, 0.5),
    ("thr0", lambda: clustered_boxes(500, 9), 0.0),
    ("thr1", lambda: clustered_boxes(500, 9), 1.0),
]


def make_pred(bs, nc, A, seed, frac=0.1, clustered=False, size=640.0, ties=False):
    """(B, 4+nc, A) xywh + class scores, fp32. `frac` of anchors get one score above 0.25."""
    r = np.random.RandomState(seed)
    pred = np.empty((bs, 4 + nc, A), F32)
    for b in range(bs):
        if clustered:
            xyxy, _ = clustered_boxes(A, seed * 100 + b, size=size)
        else:
            xyxy, _ = uniform_boxes(A, seed * 100 + b, size=size)
        cxy = (xyxy[:, :2] + xyxy[:, 2:]) / 2
        wh = xyxy[:, 2:] - xyxy[:, :2]
        pred[b, :2] = cxy.T
        pred[b, 2:4] = wh.T
        sc = r.uniform(0.0, 0.2, (nc, A)).astype(F32)
        hot = r.rand(A) < frac
        k = int(hot.sum())
        cls = r.randint(0, nc, k)
        val = r.uniform(0.26, 0.999, k).astype(F32)
        if ties:
            val = (np.round(val * 10) / 10 + 0.05).astype(F32)
        sc[cls, np.nonzero(hot)[0]] = val
        # a second, lower, above-threshold class on a few anchors (exercises multi_label)
        extra = np.nonzero(hot)[0][::7]
        sc[(cls[::7] + 1) % nc, extra] = F32(0.255)
        pred[b, 4:] = sc
    return pred


def stress_pred(bs, A=30000, nc=10, seed0=0, clustered=False):
    """BASELINE.json config 5: every anchor is a candidate, distinct scores, one class per box."""
    pred = np.zeros((bs, 4 + nc, A), F32)
    for b in range(bs):
        if clustered:
            xyxy, s = clustered_boxes(A, seed0 + b, centres=300)
        else:
            xyxy, s = uniform_boxes(A, seed0 + b)
        r = np.random.RandomState(1000 + seed0 + b)
        pred[b, :2] = ((xyxy[:, :2] + xyxy[:, 2:]) / 2).T
        pred[b, 2:4] = (xyxy[:, 2:] - xyxy[:, :2]).T
        pred[b, 4 + r.randint(0, nc, A), np.arange(A)] = s
    return pred


PRED_CASES = [
    ("rand_b2_nc10", lambda: make_pred(2, 10, 3000, 11), dict(conf_thres=0.25, iou_thres=0.7)),
    ("clustered_b3_nc10", lambda: make_pred(3, 10, 4000, 12, frac=0.4, clustered=True), dict(conf_thres=0.25, iou_thres=0.7)),
    ("clustered_iou045_maxdet20", lambda: make_pred(2, 10, 4000, 13, frac=0.5, clustered=True),
     dict(conf_thres=0.25, iou_thres=0.45, max_det=20)),
    ("no_candidates", lambda: make_pred(2, 10, 1000, 14, frac=0.0), dict(conf_thres=0.25, iou_thres=0.7)),
    ("one_image_empty", lambda: np.concatenate([make_pred(1, 10, 1000, 15), make_pred(1, 10, 1000, 16, frac=0.0)]),
     dict(conf_thres=0.25, iou_thres=0.7)),
    ("agnostic", lambda: make_pred(2, 10, 3000, 17, frac=0.4, clustered=True), dict(conf_thres=0.25, iou_thres=0.7, agnostic=True)),
    ("classes_filter", lambda: make_pred(2, 10, 3000, 18, frac=0.4, clustered=True),
     dict(conf_thres=0.25, iou_thres=0.7, classes=[1, 3, 7])),
    ("over_max_nms", lambda: make_pred(2, 10, 3000, 19, frac=0.6, clustered=True),
     dict(conf_thres=0.25, iou_thres=0.7, max_nms=500)),
    ("score_ties", lambda: make_pred(2, 10, 3000, 20, frac=0.5, clustered=True, ties=True), dict(conf_thres=0.25, iou_thres=0.6)),
    ("nc80_b1", lambda: make_pred(1, 80, 8400, 21, frac=0.2, clustered=True), dict(conf_thres=0.25, iou_thres=0.7)),
    ("nc1", lambda: make_pred(2, 1, 2000, 22, frac=0.5, clustered=True), dict(conf_thres=0.3, iou_thres=0.5)),
    ("low_conf_val_style", lambda: make_pred(1, 10, 3000, 23, frac=0.3, clustered=True), dict(conf_thres=0.001, iou_thres=0.7)),
    ("stress_small", lambda: stress_pred(2, A=4000, clustered=True), dict(conf_thres=0.25, iou_thres=0.7)),
]

# validator-style (multi_label=True): same bit-exact bar, run by test_gpu_nms.py::test_pipeline_bit_exact
PRED_CASES_MULTILABEL = [
    ("multilabel", lambda: make_pred(2, 10, 2000, 24, frac=0.4, clustered=True),
     dict(conf_thres=0.25, iou_thres=0.7, multi_label=True)),
]
