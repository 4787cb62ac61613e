"""The reference's Python surface for the inference hot path, backed by the B200 engine.

    YOLO(cfg).predict(tensor, conf=, iou=, ...)        engine/model.py:501-560  -> list of Results-like objects
    DetectionModel(cfg)(x) -> (y, [raw maps])          nn/tasks.py:129-163 / head.py:73-74
    AutoBackend(model)(im), .warmup(), .stride, ...    nn/autobackend.py:54,503-524,718-731 (the predictor's model seam)

Only what the predictor reads is mirrored. Sources: image tensors, uint8 BGR frame batches, and lists of HWC
uint8 frames (letterboxed on the device, predictor.py); files / streams / PIL belong to the reference's `ultralytics.data` loaders
and are out of scope.
"""
from types import SimpleNamespace
from typing import List, Optional

import numpy as np
import torch

from . import cfg as _cfg
from . import ops as _ops
from . import synth as _synth
from .engine import B200DetectionModel


class DetectionModel(B200DetectionModel):
    """`DetectionModel(cfg='yolov12n.yaml', ch=3, nc=None)` as in nn/tasks.py:336-379. With no `weights`, deterministic
    synthetic weights are generated (there are no checkpoints / no network); pass `weights=ref_model.state_dict()` to run a
    trained reference model."""

    def __init__(self, cfg="yolov12-sod-fusion-v5-simple", ch=3, nc=None, verbose=False, weights=None, dtype=torch.bfloat16,
                 device="cuda:0", seed=0, **kw):
        if ch != 3:
            raise NotImplementedError("only 3-channel image input is on the hot path")
        spec = _cfg.get_spec(cfg, ch=ch, nc=nc)
        name = cfg if isinstance(cfg, str) else "custom"
        name = name.rsplit("/", 1)[-1]
        name = name[:-5] if name.endswith(".yaml") else name
        name = _cfg.ALIASES.get(name, name)
        if weights is None:
            weights = _synth.synth_state_dict(spec, name, seed)
        super().__init__(spec, weights, dtype=dtype, device=device, **kw)
        self.cfg_name = name
        det = spec.layers[-1]
        # what callers read off `model.model[-1]`
        head = SimpleNamespace(nc=self.nc, nl=self.nl, reg_max=16, no=self.no, stride=self.stride, legacy=det.p["legacy"],
                               f=det.f, i=det.i, type="ultralytics.nn.modules.head.Detect")
        self.model = [SimpleNamespace(i=L.i, f=L.f, type=L.type) for L in spec.layers[:-1]] + [head]
        self.save = spec.save
        self.task = "detect"


class AutoBackend:
    """The attribute set and call signature `BasePredictor` expects from `AutoBackend` (engine/predictor.py:131,143,158-159,
    193,239,306-321). Single backend: the sm_100a library. Outputs are fresh tensors, as the reference's are."""

    def __init__(self, weights: DetectionModel, device=None, dnn=False, data=None, fp16=False, batch=1, fuse=True, verbose=False):
        if not isinstance(weights, B200DetectionModel):
            raise TypeError("AutoBackend(weights=...) must be a yolo_sod_b200 DetectionModel (in-memory module path, "
                            "autobackend.py:145-155); exported-format backends are out of scope")
        if fp16 and weights.dtype != torch.float16:
            weights = weights.half()        # autobackend.py:154 `model.half() if fp16 else model.float()`
        self.model = weights
        self.device = weights.device
        self.fp16 = weights.dtype == torch.float16
        self.stride = int(max(weights.stride_list))
        self.names = weights.names
        self.pt = True
        self.jit = self.onnx = self.engine = self.triton = self.imx = self.dynamic = self.nhwc = False
        self.task = "detect"

    def forward(self, im, augment=False, visualize=False, embed=None, **kw):
        if augment or visualize or embed is not None:
            raise NotImplementedError("augment / visualize / embed are outside the inference hot path")
        return self.model(im, **kw)

    __call__ = forward

    def eval(self):
        return self

    def warmup(self, imgsz=(1, 3, 640, 640)):
        """autobackend.py:718-731: compile + capture the program for this shape and run it once."""
        im = torch.zeros(*imgsz, dtype=torch.float32, device=self.device)
        self.forward(im)
        torch.cuda.synchronize(self.device)


class Boxes:
    """Minimal stand-in for engine/results.py `Boxes` (asserts last dim in {6,7} there, :1005): xyxy, conf, cls views."""

    def __init__(self, data: torch.Tensor, orig_shape):
        assert data.shape[-1] == 6
        self.data = data
        self.orig_shape = orig_shape

    xyxy = property(lambda s: s.data[:, :4])
    conf = property(lambda s: s.data[:, 4])
    cls = property(lambda s: s.data[:, 5])

    def __len__(self):
        return self.data.shape[0]


class Results:
    def __init__(self, boxes: torch.Tensor, orig_shape, names, speed=None):
        self.boxes = Boxes(boxes, orig_shape)
        self.orig_shape = orig_shape
        self.names = names
        self.speed = speed or {}

    def __len__(self):
        return len(self.boxes)


class ResultStream:
    """What `predict(..., stream=True)` returns: an iterator of per-image Results over a batch whose device work is in flight. `.det`
    (B, max_det, 6) and `.count` (B,) are the padded batch on the GPU (valid once the stream that produced them has run)."""

    def __init__(self, det, count, shape, names, host_count=None, ready=None):
        self.det, self.count, self.shape, self.names = det, count, shape, names
        self._host_count, self._ready = host_count, ready
        self._items = None

    def _materialise(self):
        if self._items is None:
            if self._ready is not None:
                self._ready.synchronize()      # the host sync: this batch's counts have landed in pinned memory
                n = self._host_count.tolist()
            else:
                n = self.count.tolist()
            self._items = [Results(self.det[b, :k], self.shape, self.names) for b, k in enumerate(n)]
        return self._items

    def __iter__(self):
        return iter(self._materialise())

    def __len__(self):
        return int(self.det.shape[0])


class YOLO:
    """`YOLO(cfg_yaml)` facade (models/yolo/model.py:14-23, engine/model.py:84-151,501-560) for tensor sources."""

    def __init__(self, model="yolov12-sod-fusion-v5-simple", task="detect", verbose=False, weights=None, dtype=torch.bfloat16,
                 device="cuda:0", seed=0, overlap_nms=False):
        if task not in (None, "detect"):
            raise NotImplementedError("only task='detect' is on the hot path")
        if isinstance(model, str) and model.endswith(".pt"):
            # engine/model.py:288-292 _load -> attempt_load_one_weight (nn/tasks.py:941-964): a trained reference checkpoint
            from .checkpoint import attempt_load_one_weight
            model, self.ckpt = attempt_load_one_weight(model, device=device, dtype=dtype)
        self.model = model if isinstance(model, B200DetectionModel) else DetectionModel(model, weights=weights, dtype=dtype,
                                                                                        device=device, seed=seed)
        self.backend = AutoBackend(self.model)
        self.task = "detect"
        self.names = self.model.names
        # overlap_nms: throughput mode for back-to-back batches. Batch i's NMS (sort + greedy sweep: ~0.1 ms of latency-bound, 32-CTA
        # kernels) runs on a side stream while batch i+1's forward already occupies the SMs; the forward alternates between two program
        # slots so the `y` the NMS is reading is never the one being written (measured: 3.466 -> 3.415 ms per 32-image step).
        # Results are produced on `nms_stream`: predict() / iterating a predict(stream=True) generator synchronise with it; users of
        # predict_padded() call join() (or wait on nms_stream) before touching the tensors on another stream.
        self.overlap_nms = bool(overlap_nms)
        self.nms_stream = torch.cuda.Stream(device=self.model.device) if overlap_nms else None
        self._slot = 0
        self._slot_done = [None, None]

    @torch.no_grad()
    def predict(self, source, stream=False, conf=0.25, iou=0.7, max_det=300, classes=None, agnostic_nms=False, imgsz=None,
                half=False, **kwargs) -> List[Results]:
        """Defaults follow cfg/default.yaml:51-54 and engine/model.py:547 (conf 0.25, iou 0.7, max_det 300). `half=True` runs the
        model's IEEE fp16 twin (cfg/default.yaml:60 `half`, predictor.py:309-318 -> AutoBackend(fp16=True) -> model.half())."""
        if half and not self.backend.fp16:
            if getattr(self, "_half_backend", None) is None:
                self._half_backend = AutoBackend(self.model, fp16=True)
            saved = self.backend
            self.backend = self._half_backend
            try:
                return self.predict(source, stream, conf, iou, max_det, classes, agnostic_nms, imgsz, False, **kwargs)
            finally:
                self.backend = saved
        if isinstance(source, np.ndarray):
            source = [source] if source.ndim == 3 else list(source)
        if isinstance(source, (list, tuple)):
            return self._predict_frames(list(source), conf, iou, max_det, classes, agnostic_nms, imgsz)
        if not torch.is_tensor(source):
            raise NotImplementedError("accepted sources: (B,3,H,W) float tensors in [0,1] (predictor.py:116-134 tensor branch), (B,H,W,3) "
                                      "uint8 BGR frames, or a list of HWC uint8 BGR frames (numpy / torch); file and stream sources need "
                                      "the reference's ultralytics.data loaders, which are out of scope")
        det, counts, (h, w) = self.predict_padded(source, conf, iou, max_det, classes, agnostic_nms)
        rs_stream = self.nms_stream if self.overlap_nms else torch.cuda.current_stream(counts.device)   # where det / counts are produced
        if stream:
            # engine/model.py:501-560 `stream=True`: a generator of Results instead of a list (predictor.py:197-205 stream_inference).
            # All device work of the batch is already enqueued; the host synchronisation (the per-image counts) happens when the
            # caller starts consuming the generator, so a caller may submit the next batch first.
            # the counts travel to pinned host memory right behind the NMS; consuming the generator waits for THAT copy only (an
            # event), not for whatever the caller enqueued afterwards
            ring = getattr(self, "_count_ring", None)
            if ring is None or ring[0][0].numel() < int(counts.numel()):
                ring = self._count_ring = [(torch.empty(int(counts.numel()), dtype=torch.int32).pin_memory(), torch.cuda.Event()) for _ in range(4)]
                self._count_next = 0
            hbuf, ev = ring[self._count_next % len(ring)]
            self._count_next += 1
            hview = hbuf[: int(counts.numel())]
            with torch.cuda.stream(rs_stream):
                hview.copy_(counts, non_blocking=True)
                ev.record(rs_stream)
            return ResultStream(det, counts, (h, w), self.names, hview, ev)
        if self.overlap_nms:
            self.join()
        ncount = counts.tolist()   # the one host sync of the call: the API returns variable-length per-image tensors
        return _ops.DetList([Results(det[b, :n], (h, w), self.names) for b, n in enumerate(ncount)], det, counts)

    @torch.no_grad()
    def predict_padded(self, source, conf=0.25, iou=0.7, max_det=300, classes=None, agnostic_nms=False):
        """predict() without the host synchronisation: (det (B,max_det,6) fp32 fresh tensor, count (B,) int32, (h, w)) on the GPU.
        The forward runs in the zero-copy / no-raw-map mode (its `y` is consumed by the NMS launches that follow on the same
        stream, nothing else reads it), and scale_boxes(img.shape[2:], boxes, orig_shape) with identical shapes == clip_boxes
        (ops.py:92-127,319-338) is ONE launch over the whole padded batch."""
        im = source if source.dim() == 4 else source[None]
        if im.dtype == torch.uint8:   # raw BGR HWC frames: preprocess (predictor.py:127-133) is fused into the stem kernel
            h, w = int(im.shape[1]), int(im.shape[2])
        else:
            im = im if im.dtype == torch.float32 else im.float()
            h, w = int(im.shape[2]), int(im.shape[3])
        if not self.overlap_nms:
            preds = self.backend(im, static=True, want_raw=False)
            det, count, _ = _ops.nms_padded(preds[0], conf, iou, classes=classes, agnostic=agnostic_nms, max_det=max_det)
            _ops.clip_boxes(det, (h, w))
            return det, count, (h, w)
        k = self._slot
        self._slot ^= 1
        main = torch.cuda.current_stream(self.model.device)
        if self._slot_done[k] is not None:
            main.wait_event(self._slot_done[k])      # the NMS that read this slot's `y` two batches ago has finished
        preds = self.backend(im, static=True, want_raw=False, slot=k)
        fwd = torch.cuda.Event()
        fwd.record(main)
        with torch.cuda.stream(self.nms_stream):
            self.nms_stream.wait_event(fwd)
            det, count, _ = _ops.nms_padded(preds[0], conf, iou, classes=classes, agnostic=agnostic_nms, max_det=max_det)
            _ops.clip_boxes(det, (h, w))
            done = torch.cuda.Event()
            done.record(self.nms_stream)
        self._slot_done[k] = done
        return det, count, (h, w)

    def join(self):
        """overlap_nms: makes the current stream wait for every NMS issued so far (no host synchronisation)."""
        if self.nms_stream is not None:
            torch.cuda.current_stream(self.model.device).wait_stream(self.nms_stream)

    def _predict_frames(self, frames, conf, iou, max_det, classes, agnostic_nms, imgsz):
        """The predictor's list-of-frames branch, on the device: pre_transform (LetterBox) -> preprocess fused into the stem ->
        model -> NMS -> scale_boxes/clip_boxes to each original frame (predictor.py:116-164, detect/predict.py:25-45)."""
        from . import predictor as _pred
        bad = [f for f in frames if not (getattr(f, "ndim", 0) == 3 and f.shape[2] == 3 and str(f.dtype).endswith("uint8"))]
        if bad:
            raise ValueError(f"frames must be (H,W,3) uint8 BGR arrays, got {tuple(bad[0].shape)} {bad[0].dtype}")
        imgsz = imgsz or 640
        imgsz = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        im = _pred.pre_transform(frames, imgsz, stride=self.backend.stride, device=self.model.device)
        preds = self.backend(im)
        shapes = [tuple(int(v) for v in f.shape[:2]) for f in frames]
        dets = _pred.postprocess(preds, (int(im.shape[1]), int(im.shape[2])), shapes, conf, iou, agnostic_nms, max_det, classes)
        return _ops.DetList([Results(d, s, self.names) for d, s in zip(dets, shapes)], getattr(dets, "det", None), getattr(dets, "count", None))

    __call__ = predict
