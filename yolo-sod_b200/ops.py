"""`ops.non_max_suppression` with the reference's signature (ultralytics/utils/ops.py:167-182), running the batched,
host-sync-free CUDA pipeline in csrc/nms.cu. One device->host read (the per-image counts) happens at the very end, because
the reference API returns a Python list of variable-length tensors; `nms_padded` is the sync-free variant."""
import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from . import lib as _lib

_ws_cache = {}
_params_cache = {}


def _thr_float(iou_thres: float) -> float:
    """Largest float32 <= iou_thres: `float IoU > double thr` (torchvision's CPU kernel) == `float IoU > this float`."""
    t = np.float32(iou_thres)
    if float(t) > float(iou_thres):
        t = np.nextafter(t, np.float32(-np.inf))
    return float(t)


class DetList(list):
    """The reference's return type -- a plain list of (n_i, 6) tensors -- that also keeps the padded batch its items are views of:
    `.det` (B, max_det, 6) and `.count` (B,) int32 on the GPU, for callers that move / gather the whole batch in one transfer."""

    def __init__(self, items, det, count):
        super().__init__(items)
        self.det, self.count = det, count


def nms_padded(prediction: torch.Tensor, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
               max_det=300, nc=0, max_nms=30000, max_wh=7680):
    """Returns (det (B,max_det,6) fp32 [x1,y1,x2,y2,conf,cls], count (B,) int32, index (B,max_det) int32) on the GPU with no
    host synchronisation. `index` holds, per kept box, the anchor index (or anchor*nc+cls when multi_label)."""
    _lib.require_cuda()
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if not prediction.is_cuda:
        raise _lib.YsodError("non_max_suppression: prediction must live on the GPU (no CPU fallback)")
    if prediction.dim() != 3:
        raise ValueError(f"prediction must be (B, 4+nc, A), got {tuple(prediction.shape)}")
    pred = prediction if (prediction.dtype == torch.float32 and prediction.is_contiguous()) else prediction.float().contiguous()
    B, ch, A = pred.shape
    nc = nc or (ch - 4)
    if ch - nc - 4 != 0:
        raise NotImplementedError("mask channels (nm > 0) are outside the detection hot path")
    multi_label = bool(multi_label) and nc > 1
    dev = pred.device
    cls_t = None
    n_cls = 0
    if classes is not None:
        cls_t = torch.as_tensor(list(classes), dtype=torch.int32, device=dev)
        n_cls = int(cls_t.numel())
    lib = _lib.load()
    need = int(lib.ysod_nms_workspace_bytes(B, nc, A, int(max_nms), int(multi_label)))
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    det = torch.empty((B, max_det, 6), dtype=torch.float32, device=dev)
    index = torch.empty((B, max_det), dtype=torch.int32, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("ysod_nms_batched", _lib.ptr(pred), B, nc, A, float(np.float32(conf_thres)), _thr_float(iou_thres),
                  _lib.ptr(cls_t) if cls_t is not None else None, n_cls, int(bool(agnostic)), int(multi_label), int(max_det),
                  int(max_nms), float(max_wh), _lib.ptr(det), _lib.ptr(index), _lib.ptr(count), _lib.ptr(ws), need,
                  _lib.stream_ptr())
    return det, count, index


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=(), max_det=300, nc=0, max_time_img=0.05, max_nms=30000, max_wh=7680, in_place=True,
                        rotated=False) -> List[torch.Tensor]:
    """Same arguments and return value as the reference. Differences, all deliberate:
      * the wall-clock guard (`max_time_img`, ops.py:238,312-314) is ignored -- it silently drops images;
      * `in_place` is ignored: the input tensor is never modified;
      * `labels`, `rotated=True`, mask channels and (B,N,6) end-to-end inputs raise NotImplementedError (out of the hot path).
    """
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    if rotated or (labels and len(labels)):
        raise NotImplementedError("rotated boxes / apriori labels are outside the detection hot path")
    if prediction.shape[-1] == 6:
        raise NotImplementedError("end-to-end (B,N,6) predictions are outside the detection hot path")
    det, count, _ = nms_padded(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, nc, max_nms, max_wh)
    counts = count.tolist()  # the single host sync
    return DetList([det[b, :n] for b, n in enumerate(counts)], det, count)


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float, max_keep: Optional[int] = None) -> torch.Tensor:
    """`torchvision.ops.nms(boxes, scores, iou_threshold)` on the GPU kernels: int64 indices of kept boxes, score-descending.
    `max_keep` bounds the number of keeps evaluated (the caller of ops.py:296 slices [:max_det] anyway)."""
    _lib.require_cuda()
    if not boxes.is_cuda:
        raise _lib.YsodError("nms: boxes must live on the GPU (no CPU fallback)")
    n = int(boxes.shape[0])
    dev = boxes.device
    if n == 0:
        return torch.zeros((0,), dtype=torch.int64, device=dev)
    b = boxes.float().contiguous()
    s = scores.float().contiguous()
    mk = n if max_keep is None else min(int(max_keep), n)
    lib = _lib.load()
    need = int(lib.ysod_nms_boxes_workspace_bytes(n))
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    keep = torch.empty((mk,), dtype=torch.int32, device=dev)
    nk = torch.empty((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("ysod_nms_boxes", _lib.ptr(b), _lib.ptr(s), n, _thr_float(iou_threshold), mk, _lib.ptr(keep), _lib.ptr(nk),
                  _lib.ptr(ws), need, _lib.stream_ptr())
    return keep[: int(nk.item())].long()


def xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    """ops.py:416-433 (host-side helper, torch ops; not on the hot path)."""
    assert x.shape[-1] == 4, f"input shape last dimension expected 4 but input shape is {x.shape}"
    y = torch.empty_like(x)
    wh = x[..., 2:] / 2
    y[..., :2] = x[..., :2] - wh
    y[..., 2:] = x[..., :2] + wh
    return y


def clip_boxes(boxes: torch.Tensor, shape):
    """ops.py:319-338: clamp xyxy boxes to the image (h, w), in place, on the GPU kernel (gain 1, no pad)."""
    return scale_boxes(shape, boxes, shape, ratio_pad=((1.0, 1.0), (0, 0)))


def scale_boxes(img1_shape, boxes: torch.Tensor, img0_shape, ratio_pad=None, padding=True, xywh=False) -> torch.Tensor:
    """ops.py:92-127: letterboxed-image xyxy boxes -> original-image pixels (subtract pad, divide by gain, clip). `boxes` is a
    contiguous fp32 CUDA tensor (..., k >= 4) and is modified in place like the reference's."""
    _lib.require_cuda()
    if xywh:
        raise NotImplementedError("xywh=True is not used by the detection predictor (detect/predict.py:39)")
    if not (boxes.is_cuda and boxes.dtype == torch.float32 and boxes.is_contiguous() and boxes.shape[-1] >= 4):
        raise _lib.YsodError("scale_boxes: boxes must be a contiguous fp32 CUDA tensor (..., >= 4) (no CPU fallback)")
    if boxes.numel() == 0:
        return boxes
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    else:
        gain, pad = ratio_pad[0][0], ratio_pad[1]
    if not padding:
        pad = (0, 0)
    vals = (float(np.float32(gain)), float(pad[0]), float(pad[1]), float(img0_shape[1]), float(img0_shape[0]))
    pkey = (boxes.device.index, vals)
    params = _params_cache.get(pkey)   # device-resident (gain, pad_x, pad_y, w0, h0): no per-call host->device copy
    if params is None:
        if len(_params_cache) > 256:
            _params_cache.clear()
        params = torch.tensor(vals, dtype=torch.float32).to(boxes.device)
        _params_cache[pkey] = params
    k = int(boxes.shape[-1])
    with torch.cuda.device(boxes.device):
        _lib.call("ysod_scale_boxes", _lib.ptr(boxes), 1, boxes.numel() // k, k, _lib.ptr(params), 5, _lib.stream_ptr())
    return boxes
