"""Pins oracle/model_ref.py (and yolo_sod_b200.cfg's graph resolution) against the live reference and its goldens."""
import os

import numpy as np
import pytest
import torch

import yolo_sod_b200  # noqa: F401
from yolo_sod_b200 import cfg as ycfg, synth
from oracle import model_ref, refshim
from tests.golden.make_golden_model import CASES

GOLD = os.path.join(os.path.dirname(__file__), "golden", "model_golden.npz")
PARAMS = {"yolov12-sod-fusion-v5-simple": 13_570_780, "yolov12n": 2_553_904, "yolov12m": 19_670_784}


@pytest.mark.parametrize("name", list(PARAMS))
def test_param_inventory_counts(name):
    spec = ycfg.get_spec(name)
    sh = ycfg.param_shapes(spec)
    n = sum(int(np.prod(s)) for k, s in sh.items() if "running_" not in k and "num_batches" not in k)
    assert n == PARAMS[name]


def test_strides_and_legacy_flag():
    sod = ycfg.get_spec("yolov12-sod-fusion-v5-simple")
    assert ycfg.strides_of(sod) == [4, 8, 16, 32] and sod.legacy is True and sod.nc == 10
    v12 = ycfg.get_spec("yolov12n")
    assert ycfg.strides_of(v12) == [8, 16, 32] and v12.legacy is False and v12.nc == 80


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_oracle_matches_golden(case):
    key, name, _, b, sz = case
    g = np.load(GOLD)
    spec = ycfg.get_spec(name)
    sd = synth.synth_state_dict(spec, name, 0)
    x = synth.synth_images(b, sz, seed=7)
    (y, raw), layers = model_ref.forward(spec, sd, x, ycfg.strides_of(spec), return_layers=True)
    assert np.allclose(y.numpy(), g[f"{key}_y"], rtol=1e-4, atol=2e-3)
    for l, r in enumerate(raw):
        assert np.allclose(r.numpy(), g[f"{key}_raw{l}"].astype(np.float32), rtol=2e-3, atol=2e-3)
    st = g[f"{key}_layer_stats"]
    for idx in range(len(layers) - 1):
        t = layers[idx]
        got = [float(t.mean()), float(t.std()), float(t.abs().max())]
        assert np.allclose(got, st[idx], rtol=1e-3, atol=1e-4), (idx, got, st[idx])


@pytest.mark.skipif(not refshim.available(), reason="live reference not present")
@pytest.mark.parametrize("name,ref_name", [("yolov12-sod-fusion-v5-simple", "sod"), ("yolov12n", "yolov12n"), ("yolov12m", "yolov12m"),
                                           ("E1", "E1"), ("E4", "E4"), ("E6", "E6"), ("yolov12-sod-fusion-v5-stable", "stable"), ("yolov12-sod-fusion-v5", "v5")])
def test_oracle_matches_live_reference(name, ref_name):
    spec = ycfg.get_spec(name)
    sd = synth.synth_state_dict(spec, name, 0)
    m = refshim.build(ref_name)
    ref_sd = m.state_dict()
    assert list(ref_sd) == list(sd), "state_dict names/order differ from the reference"
    assert all(tuple(ref_sd[k].shape) == tuple(sd[k].shape) for k in sd)
    m.load_state_dict(sd, strict=True)
    assert [int(s) for s in m.stride.tolist()] == ycfg.strides_of(spec)
    assert m.model[-1].legacy == spec.legacy
    x = synth.synth_images(2, 128, seed=3)
    with torch.no_grad():
        y_ref, raw_ref = m(x)
    y, raw = model_ref.forward(spec, sd, x, ycfg.strides_of(spec))
    assert torch.allclose(y, y_ref, rtol=1e-4, atol=2e-3)
    for a, b in zip(raw, raw_ref):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-4)


@pytest.mark.skipif(not refshim.available(), reason="live reference not present")
def test_builtin_graph_equals_reference_yaml():
    for name, rel in [("yolov12-sod-fusion-v5-simple", refshim.CFG["sod"]), ("yolov12n", refshim.CFG["yolov12n"]),
                      ("yolov12m", refshim.CFG["yolov12m"]), ("yolov12-sod-fusion-v5-stable", refshim.CFG["stable"]), ("yolov12-sod-fusion-v5", refshim.CFG["v5"])] \
            + [(f"E{i}", refshim.CFG[f"E{i}"]) for i in range(1, 7)]:
        a = ycfg.get_spec(name)
        b = ycfg.get_spec(os.path.join(refshim.REFERENCE_ROOT, rel))
        assert [(l.type, l.f, l.c1, l.c2, l.p) for l in a.layers] == [(l.type, l.f, l.c1, l.c2, l.p) for l in b.layers]
