"""GPU parity of the predictor glue (csrc/predictor.cu through the C ABI) against oracle/predictor_ref.py: the letterbox kernel and
scale_boxes/clip_boxes are bit-exact; `YOLO.predict(list_of_frames)` end to end equals oracle preprocess -> model -> oracle NMS ->
oracle scale_boxes."""
import numpy as np
import pytest
import torch

import yolo_sod_b200  # noqa: F401
from oracle import nms_ref, predictor_ref as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,auto,B", [((1080, 1920), True, 2), ((1080, 1920), False, 1), ((480, 640), True, 3), ((500, 333), True, 2),
                                          ((640, 640), True, 2), ((1280, 1280), False, 2), ((321, 1001), True, 1), ((97, 53), False, 2)])
def test_letterbox_bit_exact(shape, auto, B):
    from yolo_sod_b200 import predictor
    r = np.random.RandomState(shape[1])
    frames = r.randint(0, 256, (B,) + shape + (3,)).astype(np.uint8)
    want = np.stack([P.letterbox(f, (640, 640), auto, 32) for f in frames])
    got = predictor.LetterBox((640, 640), auto=auto, stride=32)(torch.from_numpy(frames).cuda())
    assert tuple(got.shape) == want.shape
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("img1,img0", [((384, 640), (1080, 1920)), ((640, 640), (500, 333)), ((640, 640), (640, 640)), ((640, 480), (2000, 1500))])
def test_scale_boxes_bit_exact(img1, img0):
    from yolo_sod_b200 import ops
    boxes = (np.random.RandomState(5).rand(777, 6) * 700 - 30).astype(np.float32)
    want = boxes.copy()
    want[:, :4] = P.scale_boxes(img1, boxes[:, :4], img0)
    got = ops.scale_boxes(img1, torch.from_numpy(boxes).cuda(), img0)
    assert np.array_equal(got.cpu().numpy(), want)          # columns 4, 5 (conf, cls) untouched
    b4 = torch.from_numpy(boxes[:, :4].copy()).cuda()
    assert np.array_equal(ops.clip_boxes(b4, img0).cpu().numpy(), P.clip_boxes(boxes[:, :4], img0))


def test_predict_list_of_frames_end_to_end():
    """Two 270x480 BGR frames -> rect letterbox (384x640) -> forward -> NMS -> boxes in original pixels."""
    from yolo_sod_b200.model import YOLO
    r = np.random.RandomState(2)
    frames = [r.randint(0, 256, (270, 480, 3)).astype(np.uint8) for _ in range(2)]
    yolo = YOLO("yolov12-sod-fusion-v5-simple", dtype=torch.bfloat16)
    res = yolo.predict(frames, conf=0.25, iou=0.7)
    x = P.preprocess(frames, (640, 640), 32)                            # oracle LetterBox + BGR->RGB + /255
    assert x.shape == (2, 3, 384, 640)
    y, _ = yolo.model(torch.from_numpy(x).cuda())                       # same kernels through the float-tensor entry
    want = nms_ref.non_max_suppression(y.cpu().numpy(), 0.25, 0.7, max_det=300)
    assert sum(len(w) for w in want) > 0
    for rr, w in zip(res, want):
        w = w.copy()
        w[:, :4] = P.scale_boxes((384, 640), w[:, :4], (270, 480))
        assert rr.orig_shape == (270, 480)
        assert np.array_equal(rr.boxes.data.cpu().numpy(), w)


def test_predict_mixed_shapes_pads_to_full_imgsz():
    from yolo_sod_b200 import predictor
    r = np.random.RandomState(4)
    frames = [r.randint(0, 256, (270, 480, 3)).astype(np.uint8), r.randint(0, 256, (600, 400, 3)).astype(np.uint8)]
    got = predictor.pre_transform(frames, (640, 640), 32)
    want = np.stack(P.pre_transform(frames, (640, 640), 32))
    assert np.array_equal(got.cpu().numpy(), want)
