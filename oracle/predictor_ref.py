"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the predictor glue either side of the forward+NMS hot path.
Nothing under yolo_sod_b200/ imports this file; it is the checker for csrc/predictor.cu.

What it restates, and where the arithmetic lives:

* `preprocess`            ultralytics/engine/predictor.py:116-134  (stack -> BGR->RGB -> BHWC->BCHW -> float -> /255)
* `pre_transform`         ultralytics/engine/predictor.py:145-164  (LetterBox(imgsz, auto=same_shapes and pt, stride))
* `LetterBox.__call__`    ultralytics/data/augment.py of upstream ultralytics 8.3.63 -- the reference checkout does NOT contain
                          `ultralytics/data` (.gitignore:11), so this follows the published upstream algorithm:
                          r = min(new/old), new_unpad = round(old*r), minimum-rectangle padding mod stride when auto, resize with
                          cv2.INTER_LINEAR, copyMakeBorder with 114 using round(d -/+ 0.1).
* `resize_linear_u8`      cv2.resize(..., interpolation=cv2.INTER_LINEAR) for uint8 (OpenCV 4.x imgproc/resize.cpp: 11-bit fixed
                          point coefficients, HResizeLinear / VResizeLinear with FixedPtCast, and the 2x2 -> INTER_AREA shortcut).
                          Third-party (opencv-python, requirements.txt of the reference pins opencv-python>=4.6.0); pinned here
                          against the installed cv2 4.13 by tests/test_oracle_predictor.py.
* `scale_boxes`, `clip_boxes`   ultralytics/utils/ops.py:92-127, 319-338 -- pinned against the live reference functions.
"""
import math

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _py_round(x):
    return int(round(x))      # Python round: half to even, as the reference's int(round(...))


def letterbox_geometry(shape, new_shape=(640, 640), auto=False, stride=32, scaleup=True, center=True):
    """LetterBox geometry: returns dict(new_unpad=(w,h), top, bottom, left, right, out_shape=(H,W), ratio=r)."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    h0, w0 = int(shape[0]), int(shape[1])
    r = min(new_shape[0] / h0, new_shape[1] / w0)
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = (_py_round(w0 * r), _py_round(h0 * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = dw % stride, dh % stride      # np.mod on non-negative ints
    if center:
        dw /= 2
        dh /= 2
    top, bottom = (_py_round(dh - 0.1) if center else 0), _py_round(dh + 0.1)
    left, right = (_py_round(dw - 0.1) if center else 0), _py_round(dw + 0.1)
    return dict(new_unpad=new_unpad, top=top, bottom=bottom, left=left, right=right,
                out_shape=(new_unpad[1] + top + bottom, new_unpad[0] + left + right), ratio=r)


def _cv_round_short(v):
    """saturate_cast<short>(float): cvRound = round half to even, on a float32 value."""
    return np.clip(np.rint(v.astype(np.float32)), -32768, 32767).astype(np.int32)


def linear_tables(src, dst):
    """Per destination index: source index, clamped flag and the two 11-bit coefficients, as cv::resize builds them
    (resize.cpp: fx = (float)((dx + 0.5) * scale - 0.5); sx = floor(fx); fx -= sx; border handling for ksize = 2)."""
    scale = 1.0 / (dst / src)                      # scale_x = 1. / inv_scale_x, inv_scale_x = (double)dsize / ssize
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def resize_linear_u8(img, new_wh):
    """cv2.resize(img, new_wh, interpolation=cv2.INTER_LINEAR) for (H,W,C) uint8, bit for bit."""
    h0, w0 = img.shape[:2]
    w1, h1 = int(new_wh[0]), int(new_wh[1])
    src = img.astype(np.int32)
    if w0 == 2 * w1 and h0 == 2 * h1:
        # "in case of scale_x && scale_y is equal to 2, INTER_AREA (fast) also is equal to INTER_LINEAR": ResizeAreaFast 2x2
        s = src[0::2, 0::2] + src[0::2, 1::2] + src[1::2, 0::2] + src[1::2, 1::2]
        return ((s + 2) >> 2).astype(np.uint8)
    sx, fx = linear_tables(w0, w1)
    sy, fy = linear_tables(h0, h1)
    # horizontal: sx < 0 -> (0, fx = 0); sx >= w0 - 1 -> (w0 - 1, fx = 0)
    lo = sx < 0
    hi = sx >= w0 - 1
    fx = np.where(lo | hi, np.float32(0), fx).astype(np.float32)
    sx = np.where(lo, 0, np.where(hi, w0 - 1, sx))
    a0 = _cv_round_short((np.float32(1.0) - fx) * np.float32(COEF_SCALE))
    a1 = _cv_round_short(fx * np.float32(COEF_SCALE))
    sx1 = np.minimum(sx + 1, w0 - 1)
    a1 = np.where(hi, 0, a1)                       # dx >= xmax: D = S[sx] * ONE
    a0 = np.where(hi, COEF_SCALE, a0)
    # vertical: coefficients are kept, the two row indices are clamped to [0, h0 - 1]
    b0 = _cv_round_short((np.float32(1.0) - fy) * np.float32(COEF_SCALE))
    b1 = _cv_round_short(fy * np.float32(COEF_SCALE))
    r0 = np.clip(sy, 0, h0 - 1)
    r1 = np.clip(sy + 1, 0, h0 - 1)
    rows = src[:, sx] * a0[None, :, None] + src[:, sx1] * a1[None, :, None]          # (h0, w1, C) int32, scaled by 2^11
    s0, s1 = rows[r0], rows[r1]
    out = (((b0[:, None, None] * (s0 >> 4)) >> 16) + ((b1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox(img, new_shape=(640, 640), auto=False, stride=32, value=114):
    """LetterBox(new_shape, auto, stride=stride)(image=img) for an (H,W,3) uint8 BGR frame."""
    g = letterbox_geometry(img.shape[:2], new_shape, auto, stride)
    if (img.shape[1], img.shape[0]) != g["new_unpad"]:
        img = resize_linear_u8(img, g["new_unpad"])
    H, W = g["out_shape"]
    out = np.full((H, W, img.shape[2]), value, dtype=np.uint8)
    out[g["top"]:g["top"] + img.shape[0], g["left"]:g["left"] + img.shape[1]] = img
    return out


def pre_transform(frames, imgsz=(640, 640), stride=32, pt=True):
    """predictor.py:145-164: LetterBox(imgsz, auto=same_shapes and pt, stride) on every frame."""
    same = len({f.shape for f in frames}) == 1
    return [letterbox(f, imgsz, auto=same and pt, stride=stride) for f in frames]


def preprocess(frames, imgsz=(640, 640), stride=32):
    """predictor.py:116-134 for a list of HWC BGR uint8 frames: (B,3,H,W) float32 in [0,1]."""
    im = np.stack(pre_transform(frames, imgsz, stride))
    im = np.ascontiguousarray(im[..., ::-1].transpose((0, 3, 1, 2)))
    return im.astype(np.float32) / np.float32(255)


def clip_boxes(boxes, shape):
    """ops.py:319-338 (torch branch): clamp x to [0, w], y to [0, h]."""
    boxes = boxes.copy()
    boxes[..., 0] = np.clip(boxes[..., 0], 0, shape[1])
    boxes[..., 1] = np.clip(boxes[..., 1], 0, shape[0])
    boxes[..., 2] = np.clip(boxes[..., 2], 0, shape[1])
    boxes[..., 3] = np.clip(boxes[..., 3], 0, shape[0])
    return boxes


def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None, padding=True):
    """ops.py:92-127 for xyxy float32 boxes (n,4): subtract the letterbox pad, divide by the gain, clip to the original image."""
    boxes = np.asarray(boxes, dtype=np.float32).copy()
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (_py_round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), _py_round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    if padding:
        boxes[..., 0] -= np.float32(pad[0])
        boxes[..., 1] -= np.float32(pad[1])
        boxes[..., 2] -= np.float32(pad[0])
        boxes[..., 3] -= np.float32(pad[1])
    boxes[..., :4] = boxes[..., :4] / np.float32(gain)     # torch: fp32 tensor /= python float -> fp32 true division (CPU)
    return clip_boxes(boxes, img0_shape)


def scale_params(img1_shape, img0_shape):
    """(gain as float32, pad_x, pad_y) of ops.py:111-116 -- the three numbers the fused kernel takes per image."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    return (float(np.float32(gain)), _py_round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1),
            _py_round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
