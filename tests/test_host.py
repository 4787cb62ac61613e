"""CPU tests of the host side: C-ABI export table, graph resolution, synthetic weights, failure behaviour without a GPU."""
import ctypes
import os
import re

import pytest
import torch

import yolo_sod_b200  # noqa: F401
from yolo_sod_b200 import cfg as ycfg, lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    if not (os.path.exists(lib.LIB_PATH) and os.path.exists(lib.LIB_PATH_F16)):
        import __graft_entry__ as g
        g.build()


def test_library_exports_every_declared_symbol():
    _built()
    header = open(os.path.join(ROOT, "include", "ysod.h")).read()
    declared = set(re.findall(r"\b(ysod_[a-z0-9_]+)\s*\(", header))
    declared.discard("ysod_conv_tc")
    assert len(declared) >= 30
    for path in (lib.LIB_PATH, lib.LIB_PATH_F16):       # the bf16 build and the fp16 (`half=True`) build export the same ABI
        so = ctypes.CDLL(path)
        for name in sorted(declared):
            assert hasattr(so, name), f"{name} declared in include/ysod.h but not exported by {os.path.basename(path)}"
    assert declared == set(lib.PROTOTYPES), "lib.py prototypes and include/ysod.h disagree"
    assert lib.load().ysod_version() == 100 and lib.load().ysod_compiled_arch() == 100
    assert lib.load().ysod_storage_dtype() == 1 and lib.load(half=True).ysod_storage_dtype() == 2


def test_binding_prototypes_have_the_header_arity():
    """Every ctypes prototype in lib.py passes as many arguments as include/ysod.h declares (catches a drifted binding before a GPU run)."""
    header = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "ysod.h")).read(), flags=re.S)
    decls = dict(re.findall(r"\b(ysod_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S))
    decls.pop("ysod_conv_tc", None)
    checked = 0
    for name, (res, args) in lib.PROTOTYPES.items():
        assert name in decls, name
        params = decls[name].strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), f"{name}: header declares {n} parameters, lib.py passes {len(args)}"
        checked += 1
    assert checked >= 30


def test_error_reporting_without_compute():
    _built()
    so = lib.load()
    rc = so.ysod_dfl_decode(None, 0, 1, 1, 1, 80, 10, 16, 4.0, None, 1, 0, None)
    assert rc != 0 and b"null" in so.ysod_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from yolo_sod_b200 import ops
    from yolo_sod_b200.model import DetectionModel
    with pytest.raises(lib.YsodError):
        ops.non_max_suppression(torch.zeros(1, 14, 10), 0.25, 0.7)
    with pytest.raises(lib.YsodError):
        DetectionModel("yolov12n")


def test_thr_float_rounding():
    from yolo_sod_b200.ops import _thr_float
    import numpy as np
    for t in [0.0, 0.3, 0.45, 0.5, 0.6, 0.7, 0.95, 1.0]:
        f = np.float32(_thr_float(t))
        assert float(f) <= t and float(np.nextafter(f, np.float32(2))) > t


def test_graph_resolution_sod():
    spec = ycfg.get_spec("yolov12-sod-fusion-v5-simple")
    types = [l.type for l in spec.layers]
    assert len(types) == 40 and types[1] == "SE_Block" and types[9] == "SwinBlock" and types[12] == "A2_Attn" and types[-1] == "Detect"
    assert [l.c2 for l in spec.layers[:14]] == [32, 32, 64, 64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512]
    assert spec.layers[1].p["reduction"] == 64 and spec.layers[23].p["reduction"] == 256      # SE_Block arg is the reduction
    assert spec.layers[3].p["n"] == 1 and spec.layers[6].p["n"] == 2                           # depth gain 0.33
    assert spec.layers[-1].c1 == [64, 128, 256, 512] and spec.layers[-1].p["legacy"] is True
    assert spec.save == sorted({3, 6, 9, 13, 18, 23, 28, 32, 35, 38})


def test_graph_resolution_yolov12_scales():
    n, m = ycfg.get_spec("yolov12n"), ycfg.get_spec("yolov12m.yaml")
    assert [l.c2 for l in n.layers[:9]] == [16, 32, 64, 64, 128, 128, 128, 256, 256]
    assert [l.c2 for l in m.layers[:9]] == [64, 128, 256, 256, 512, 512, 512, 512, 512]
    assert n.layers[2].p["c3k"] is False and m.layers[2].p["c3k"] is True
    assert n.layers[6].p == dict(n=2, a2=True, area=4, residual=False, mlp_ratio=2.0, e=0.5, g=1, shortcut=True)
    assert n.layers[1].p["g"] == 2 and n.layers[3].p["g"] == 4
    with pytest.raises(NotImplementedError):
        ycfg.parse_model({"nc": 1, "backbone": [[-1, 1, "Focus", [64, 3]]], "head": []})


def test_synth_weights_deterministic_and_complete():
    spec = ycfg.get_spec("yolov12n")
    a = synth.synth_state_dict(spec, "yolov12n", 0)
    b = synth.synth_state_dict(spec, "yolov12n", 0)
    c = synth.synth_state_dict(spec, "yolov12n", 1)
    assert list(a) == list(ycfg.param_shapes(spec))
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert any(not torch.equal(a[k], c[k]) for k in a)
    assert all(tuple(a[k].shape) == tuple(s) for k, s in ycfg.param_shapes(spec).items())
