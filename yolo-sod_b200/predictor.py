"""Predictor glue on the device: what `BasePredictor` does either side of the model call,
for frames that are already in GPU (or pinned host) memory.

    LetterBox(new_shape, auto, stride)(frames)   ultralytics/data/augment.py of upstream 8.3.63 (absent from the reference checkout,
                                                 .gitignore:11) as used by engine/predictor.py:145-164 pre_transform
    DetectionPredictor.preprocess / postprocess  engine/predictor.py:116-134, models/yolo/detect/predict.py:25-45

Host code only computes geometry (a handful of integers per distinct frame shape); pixels and boxes are touched by
csrc/predictor.cu (letterbox, scale_boxes+clip_boxes), csrc/stem.cu (BGR->RGB, HWC->CHW, /255 fused into the stem) and csrc/nms.cu.
"""
from typing import List, Sequence

import numpy as np
import torch

from . import lib as _lib
from . import ops as _ops


def letterbox_geometry(shape, new_shape=(640, 640), auto=False, stride=32, scaleup=True, center=True):
    """LetterBox.__call__ geometry (upstream augment.py): r = min(new/old); new_unpad = round(old * r); padding = remainder
    (mod stride when `auto`), split with round(d -/+ 0.1). Returns dict(new_unpad=(w,h), top, bottom, left, right, out_shape=(H,W))."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    h0, w0 = int(shape[0]), int(shape[1])
    r = min(new_shape[0] / h0, new_shape[1] / w0)
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = (int(round(w0 * r)), int(round(h0 * r)))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = dw % stride, dh % stride
    if center:
        dw /= 2
        dh /= 2
    top, bottom = (int(round(dh - 0.1)) if center else 0), int(round(dh + 0.1))
    left, right = (int(round(dw - 0.1)) if center else 0), int(round(dw + 0.1))
    return dict(new_unpad=new_unpad, top=top, bottom=bottom, left=left, right=right,
                out_shape=(new_unpad[1] + top + bottom, new_unpad[0] + left + right))


class LetterBox:
    """`LetterBox(new_shape, auto=False, stride=32)` for batches of equally-sized uint8 BGR frames on the GPU:
    `(B,H0,W0,3) uint8 cuda -> (B,H,W,3) uint8 cuda`, bit-identical to cv2.resize(INTER_LINEAR) + cv2.copyMakeBorder(114)."""

    def __init__(self, new_shape=(640, 640), auto=False, scaleFill=False, scaleup=True, center=True, stride=32):
        if scaleFill:
            raise NotImplementedError("scaleFill is not used by the predictor (predictor.py:158-163)")
        self.new_shape = (new_shape, new_shape) if isinstance(new_shape, int) else tuple(new_shape)
        self.auto, self.scaleup, self.center, self.stride = auto, scaleup, center, stride

    def geometry(self, shape):
        return letterbox_geometry(shape, self.new_shape, self.auto, self.stride, self.scaleup, self.center)

    def __call__(self, frames: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        _lib.require_cuda()
        if not (torch.is_tensor(frames) and frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3):
            raise _lib.YsodError("LetterBox: frames must be a (B,H0,W0,3) uint8 CUDA tensor (no CPU fallback)")
        frames = frames.contiguous()
        B, H0, W0 = int(frames.shape[0]), int(frames.shape[1]), int(frames.shape[2])
        g = self.geometry((H0, W0))
        H, W = g["out_shape"]
        if out is None:
            out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=frames.device)
        assert tuple(out.shape) == (B, H, W, 3) and out.is_contiguous()
        with torch.cuda.device(frames.device):
            _lib.call("ysod_letterbox_u8", _lib.ptr(frames), B, H0, W0, _lib.ptr(out), H, W, g["new_unpad"][1], g["new_unpad"][0],
                      g["top"], g["left"], 114, _lib.stream_ptr())
        return out


def scale_params(img1_shape, img0_shape):
    """(gain, pad_x, pad_y, w0, h0) of ops.py:111-116 for one image."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    return (float(np.float32(gain)), float(round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1)),
            float(round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)), float(img0_shape[1]), float(img0_shape[0]))


def pre_transform(frames: Sequence, imgsz=(640, 640), stride=32, device="cuda:0") -> torch.Tensor:
    """predictor.py:145-164 for a list of HWC BGR uint8 frames (numpy arrays, or torch uint8 tensors on any device):
    LetterBox(imgsz, auto=same_shapes (pt model), stride) -> one (B,H,W,3) uint8 CUDA tensor."""
    shapes = [tuple(int(v) for v in f.shape) for f in frames]
    same = len(set(shapes)) == 1
    lb = LetterBox(imgsz, auto=same, stride=stride)
    dev = torch.device(device)

    def to_dev(f):
        t = f if torch.is_tensor(f) else torch.from_numpy(np.ascontiguousarray(f))
        return t.to(dev, non_blocking=True)

    if same:
        return lb(torch.stack([to_dev(f) for f in frames]))
    H, W = lb.new_shape                         # mixed shapes: every frame is padded to the full imgsz (auto=False)
    out = torch.empty((len(frames), H, W, 3), dtype=torch.uint8, device=dev)
    for i, f in enumerate(frames):
        lb(to_dev(f)[None], out=out[i:i + 1])
    return out


def postprocess(preds, img_shape, orig_shapes: List[tuple], conf=0.25, iou=0.7, agnostic=False, max_det=300, classes=None):
    """detect/predict.py:25-45: NMS, then scale_boxes(img.shape[2:], pred[:, :4], orig_img.shape) per image -- here one batched
    kernel on the padded NMS output. Returns a list of (n_i, 6) tensors in original-image pixels."""
    det, count, _ = _ops.nms_padded(preds[0] if isinstance(preds, (list, tuple)) else preds, conf, iou, classes=classes, agnostic=agnostic,
                                    max_det=max_det)
    params = torch.tensor([scale_params(img_shape, s[:2]) for s in orig_shapes], dtype=torch.float32).to(det.device, non_blocking=True)
    with torch.cuda.device(det.device):
        _lib.call("ysod_scale_boxes", _lib.ptr(det), int(det.shape[0]), int(det.shape[1]), 6, _lib.ptr(params), 5, _lib.stream_ptr())
    counts = count.tolist()
    return _ops.DetList([det[b, :n] for b, n in enumerate(counts)], det, count)
