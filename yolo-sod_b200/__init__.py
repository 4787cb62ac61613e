"""yolo-sod B200: hand-written sm_100a inference path (forward + decode + NMS) for the YOLOv12-SOD detector family,
behind the reference's Python surface (DetectionModel.forward -> (y, raw), ops.non_max_suppression, YOLO(cfg).predict)."""
from . import cfg  # noqa: F401

__version__ = "0.1.0"
