"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's NMS path (parity oracle).

Not part of the product. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this file. The product path (yolo-sod_b200) never does.

What is restated, and from where (paths relative to the reference checkout):

* `non_max_suppression`      ultralytics/utils/ops.py:167-316  (multi_label=False/True, labels=(),
                             rotated=False, nm=0 branch)
* `xywh2xyxy`                ultralytics/utils/ops.py:416-433
* `nms`                      torchvision.ops.nms  -- third-party, NOT under /root/reference.
                             Pinned by the reference at torchvision==0.20.1 (requirements.txt:62),
                             call site ultralytics/utils/ops.py:296. The published algorithm
                             (torchvision/csrc/ops/cpu/nms_kernel.cpp, nms_kernel_impl) is restated:
                             stable descending sort on score, greedy sweep, suppress j when
                             inter / (area_i + area_j - inter) > iou_threshold, all in fp32,
                             the comparison against the *double* threshold.

Parity status: PINNED -- tests/test_oracle_nms.py checks `nms` against the live
torchvision.ops.nms (installed in the image, 0.26.0) on random, clustered, tied, degenerate and
NaN inputs, and `non_max_suppression` against the live reference function imported through
oracle/refshim.py (where /root/reference exists) and against committed goldens in tests/golden/.

Everything is numpy float32; every arithmetic step is a separately rounded fp32 operation
(numpy never contracts to FMA), which is what the reference's C++ does on x86-64.
"""
import ctypes
import os

import numpy as np

F32 = np.float32

_CLIB = None


def _clib():
    """The C restatement (oracle/nms_ref.c), if `make -C oracle` has been run. Same algorithm,
    ~100x faster than the numpy loop for 30k boxes; used for the full-size parity cases and as
    the CPU baseline in bench.py."""
    global _CLIB
    if _CLIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libysod_oracle.so")
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.ysod_ref_nms.restype = ctypes.c_int64
            lib.ysod_ref_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double,
                                         ctypes.c_int64, ctypes.c_void_p]
            _CLIB = lib
        else:
            _CLIB = False
    return _CLIB


def nms_c(boxes, scores, iou_threshold, limit=-1):
    lib = _clib()
    if not lib:
        raise RuntimeError("oracle/_build/libysod_oracle.so missing: run `make -C oracle`")
    boxes = np.ascontiguousarray(boxes, dtype=F32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=F32).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    nk = lib.ysod_ref_nms(boxes.ctypes.data, scores.ctypes.data, n, float(iou_threshold), int(limit),
                          keep.ctypes.data)
    return keep[:nk].copy()


def xywh2xyxy(x: np.ndarray) -> np.ndarray:
    """ops.py:416-433: wh = x[..., 2:] / 2 ; xy - wh ; xy + wh."""
    x = np.asarray(x, dtype=F32)
    y = np.empty_like(x)
    xy = x[..., :2]
    wh = x[..., 2:] / F32(2)
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def nms(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float, limit: int = -1) -> np.ndarray:
    """torchvision.ops.nms (CPU kernel) restated. Returns int64 indices, score-descending.

    `limit` > 0 stops after that many keeps (the caller slices [:max_det] anyway, ops.py:297);
    it does not change the first `limit` entries.
    """
    boxes = np.ascontiguousarray(boxes, dtype=F32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=F32).reshape(-1)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)  # fp32, two roundings
    # stable descending sort: equal scores keep ascending original index
    order = np.argsort(-scores.astype(np.float64), kind="stable") if not np.isnan(scores).any() else _nan_sort(scores)
    thr = float(iou_threshold)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    zero = F32(0)
    with np.errstate(all="ignore"):
        for _i in range(n):
            i = order[_i]
            if suppressed[i]:
                continue
            keep.append(i)
            if 0 < limit <= len(keep):
                break
            rest = order[_i + 1:]
            if rest.size == 0:
                break
            xx1 = np.maximum(x1[i], x1[rest])
            yy1 = np.maximum(y1[i], y1[rest])
            xx2 = np.minimum(x2[i], x2[rest])
            yy2 = np.minimum(y2[i], y2[rest])
            w = np.maximum(zero, xx2 - xx1)
            h = np.maximum(zero, yy2 - yy1)
            inter = w * h
            ovr = inter / (areas[i] + areas[rest] - inter)
            suppressed[rest[ovr.astype(np.float64) > thr]] = True  # NaN > thr is False -> kept
    return np.asarray(keep, dtype=np.int64)


def _nan_sort(scores):
    # torch.sort(descending=True) places NaN first; keep that corner identical.
    idx = np.arange(scores.shape[0])
    isn = np.isnan(scores)
    a = idx[isn]
    b = idx[~isn]
    b = b[np.argsort(-scores[b].astype(np.float64), kind="stable")]
    return np.concatenate([a, b])


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, labels=(), max_det=300, nc=0, max_time_img=0.05,
                        max_nms=30000, max_wh=7680, in_place=True, rotated=False, return_index=False):
    """ops.py:167-316 restated for the detection hot path (nm == 0, labels == (), not rotated).

    prediction: (B, 4+nc, A) float32, xywh + class scores. Returns a list of (n_i, 6) float32
    arrays [x1, y1, x2, y2, conf, cls]. The wall-clock guard (ops.py:238,312-314) is NOT
    restated: it truncates results nondeterministically.
    """
    assert 0 <= conf_thres <= 1
    assert 0 <= iou_thres <= 1
    assert not rotated and not labels
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    prediction = np.asarray(prediction, dtype=F32)
    bs = prediction.shape[0]
    nc = nc or (prediction.shape[1] - 4)
    assert prediction.shape[1] - nc - 4 == 0, "mask channels (nm>0) are outside the hot path"
    conf_t = F32(conf_thres)  # torch compares fp32 tensor > python float as fp32? see note below
    # torch: `tensor > python_float` converts the scalar to the tensor dtype only for the
    # comparison kernel's opmath; for float32 tensors the scalar is cast to float32 -> use F32.
    xc = prediction[:, 4:4 + nc].max(1) > conf_t
    pred = np.transpose(prediction, (0, 2, 1))  # (B, A, 4+nc)
    out = [np.zeros((0, 6), dtype=F32) for _ in range(bs)]
    out_idx = [np.zeros((0,), dtype=np.int64) for _ in range(bs)]  # candidate id: anchor, or anchor*nc+cls if multi_label
    for xi in range(bs):
        x = pred[xi][xc[xi]]
        ids = np.nonzero(xc[xi])[0]
        if not x.shape[0]:
            continue
        box = xywh2xyxy(x[:, :4])
        cls = x[:, 4:4 + nc]
        if multi_label and nc > 1:
            i, j = np.nonzero(cls > conf_t)
            x = np.concatenate((box[i], cls[i, j][:, None], j[:, None].astype(F32)), 1)
            ids = ids[i] * nc + j
        else:
            j = cls.argmax(1)  # first maximal index, as torch.max(dim)
            conf = cls[np.arange(cls.shape[0]), j]
            x = np.concatenate((box, conf[:, None], j[:, None].astype(F32)), 1)[conf > conf_t]
            ids = ids[conf > conf_t]
        if classes is not None:
            sel = np.isin(x[:, 5], np.asarray(classes, dtype=F32))
            x, ids = x[sel], ids[sel]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:
            # reference: unstable argsort(descending); ties are implementation-defined there.
            # restated as stable (lower index first), which is what torch's CPU sort yields.
            top = np.argsort(-x[:, 4].astype(np.float64), kind="stable")[:max_nms]
            x, ids = x[top], ids[top]
        c = x[:, 5:6] * F32(0 if agnostic else max_wh)
        boxes = x[:, :4] + c
        keep = (nms_c if (_clib() and not np.isnan(x[:, 4]).any()) else nms)(boxes, x[:, 4], iou_thres, limit=max_det)[:max_det]
        out[xi] = x[keep]
        out_idx[xi] = ids[keep]
    return (out, out_idx) if return_index else out
