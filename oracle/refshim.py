"""TEST INFRASTRUCTURE ONLY -- import shim for the *live* reference (quitedob/yolo-sod).

Only usable where /root/reference exists (the build container). Nothing in the product
path, `-m gpu` tests, smoke() or bench.py may import this module: the reference does not
travel to the GPU box. It is used by tests/golden/make_golden.py (fixture generation) and
by the CPU tests that pin oracle/ against the real reference code.

Recipe follows SURVEY.md section 8(c) / Appendix B: stub the two missing imports
(matplotlib, thop), and register a bare `ultralytics` package object so that
ultralytics/__init__.py (which needs the un-committed ultralytics.data) never executes.
"""
import importlib.machinery as _M
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("YSOD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ultralytics", "nn"))


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Stub(f"{self.__name__}.{k}")
        setattr(self, k, m)
        return m

    def __call__(self, *a, **k):
        return None


_loaded = None


def load():
    """Returns (DetectionModel, ops_module, tasks_module) of the live reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference checkout not found under {REFERENCE_ROOT}")
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.image", "thop"):
        if n not in sys.modules:
            m = _Stub(n)
            m.__path__ = []
            m.__spec__ = _M.ModuleSpec(n, None)
            sys.modules[n] = m
    if "ultralytics" not in sys.modules:
        pkg = types.ModuleType("ultralytics")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "ultralytics")]
        pkg.__version__ = "8.3.63"
        pkg.__spec__ = _M.ModuleSpec("ultralytics", None, is_package=True)
        pkg.__spec__.submodule_search_locations = pkg.__path__
        sys.modules["ultralytics"] = pkg
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/ysod_cfg")
    os.makedirs(os.environ["YOLO_CONFIG_DIR"], exist_ok=True)
    import logging
    logging.disable(logging.WARNING)
    from ultralytics.nn import tasks
    from ultralytics.utils import ops
    logging.disable(logging.NOTSET)
    _loaded = (tasks.DetectionModel, ops, tasks)
    return _loaded


CFG = {
    "sod": "ultralytics/cfg/models/new/yolov12-sod-fusion-v5-simple.yaml",
    "stable": "ultralytics/cfg/models/new/yolov12-sod-fusion-v5-stable.yaml",
    "v5": "ultralytics/cfg/models/new/yolov12-sod-fusion-v5.yaml",
    "yolov12n": "ultralytics/cfg/models/v12/yolov12n.yaml",
    "yolov12s": "ultralytics/cfg/models/v12/yolov12s.yaml",
    "yolov12m": "ultralytics/cfg/models/v12/yolov12m.yaml",
    **{f"E{i}": f"ultralytics/cfg/models/new/E{i}.yaml" for i in range(1, 7)},   # the ablation ladder (README.md:131-137)
}


def build(name: str):
    """Builds the live reference DetectionModel for one of the named configs (eval mode)."""
    DetectionModel, _, _ = load()
    return DetectionModel(os.path.join(REFERENCE_ROOT, CFG[name]), verbose=False).eval()
