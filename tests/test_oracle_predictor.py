"""CPU: pins the predictor-glue oracle (oracle/predictor_ref.py) -- the cv2.INTER_LINEAR restatement against the installed cv2,
LetterBox against cv2.resize + cv2.copyMakeBorder with the upstream geometry, scale_boxes / clip_boxes against the live reference
(ultralytics/utils/ops.py:92-127,319-338) -- and the host-side geometry of yolo_sod_b200.predictor against the oracle."""
import numpy as np
import pytest

import yolo_sod_b200  # noqa: F401
from oracle import predictor_ref as P, refshim

cv2 = pytest.importorskip("cv2")

RESIZE_CASES = [((1080, 1920), (640, 360)), ((1080, 1920), (384, 216)), ((480, 640), (640, 480)), ((333, 517), (640, 412)),
                ((1280, 1280), (640, 640)), ((100, 100), (640, 640)), ((37, 53), (200, 100)), ((1279, 1281), (640, 640)),
                ((9, 700), (350, 5)), ((2160, 3840), (640, 360)),
                ((1, 1), (5, 7)), ((2, 2), (9, 3)), ((3, 3), (640, 640)), ((17, 16), (8, 8)), ((16, 16), (8, 8)), ((4000, 3), (2, 300))]


@pytest.mark.parametrize("case", RESIZE_CASES, ids=lambda c: f"{c[0][0]}x{c[0][1]}to{c[1][1]}x{c[1][0]}")
def test_resize_restatement_equals_cv2(case):
    (h0, w0), (w1, h1) = case
    img = np.random.RandomState(h0 + w1).randint(0, 256, (h0, w0, 3)).astype(np.uint8)
    assert np.array_equal(cv2.resize(img, (w1, h1), interpolation=cv2.INTER_LINEAR), P.resize_linear_u8(img, (w1, h1)))


def test_resize_random_shapes_equal_cv2():
    r = np.random.RandomState(7)
    for _ in range(25):
        h0, w0 = r.randint(8, 700, 2)
        w1, h1 = r.randint(8, 600, 2)
        img = r.randint(0, 256, (h0, w0, 3)).astype(np.uint8)
        assert np.array_equal(cv2.resize(img, (int(w1), int(h1)), interpolation=cv2.INTER_LINEAR), P.resize_linear_u8(img, (int(w1), int(h1)))), (h0, w0, w1, h1)


@pytest.mark.parametrize("shape,auto", [((1080, 1920), True), ((1080, 1920), False), ((480, 640), True), ((500, 333), True),
                                        ((640, 640), True), ((1280, 1280), False), ((321, 1001), True)])
def test_letterbox_equals_cv2_pipeline(shape, auto):
    img = np.random.RandomState(shape[0]).randint(0, 256, shape + (3,)).astype(np.uint8)
    g = P.letterbox_geometry(shape, (640, 640), auto, 32)
    ref = img
    if (shape[1], shape[0]) != g["new_unpad"]:
        ref = cv2.resize(ref, g["new_unpad"], interpolation=cv2.INTER_LINEAR)
    ref = cv2.copyMakeBorder(ref, g["top"], g["bottom"], g["left"], g["right"], cv2.BORDER_CONSTANT, value=(114, 114, 114))
    got = P.letterbox(img, (640, 640), auto, 32)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    assert got.shape[0] % 32 == 0 and got.shape[1] % 32 == 0 or not auto
    from yolo_sod_b200 import predictor
    assert predictor.letterbox_geometry(shape, (640, 640), auto, 32) == {k: v for k, v in g.items() if k != "ratio"}


@pytest.mark.skipif(not refshim.available(), reason="live reference not present")
@pytest.mark.parametrize("img1,img0", [((384, 640), (1080, 1920)), ((640, 640), (500, 333)), ((640, 640), (640, 640)), ((640, 480), (2000, 1500))])
def test_scale_boxes_equals_live_reference(img1, img0):
    import torch
    ops = refshim.load()[1]
    r = np.random.RandomState(3)
    boxes = (r.rand(500, 4) * 700 - 30).astype(np.float32)
    want = ops.scale_boxes(img1, torch.from_numpy(boxes.copy()), img0).numpy()
    got = P.scale_boxes(img1, boxes, img0)
    assert np.array_equal(want, got)
    gain, px, py = P.scale_params(img1, img0)
    from yolo_sod_b200 import predictor
    assert predictor.scale_params(img1, img0) == (gain, float(px), float(py), float(img0[1]), float(img0[0]))
    assert np.array_equal(ops.clip_boxes(torch.from_numpy(boxes.copy()), img0).numpy(), P.clip_boxes(boxes, img0))
