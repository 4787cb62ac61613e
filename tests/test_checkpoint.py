"""Checkpoint ingestion (yolo_sod_b200/checkpoint.py, mirrors nn/tasks.py:860-964): a reference-format `.pt` written by the live
reference (tests/golden/make_golden_ckpt.py) is read without the reference package; the GPU test compiles it and checks the forward."""
import io
import os
import pickle
import sys

import pytest
import torch

import yolo_sod_b200  # noqa: F401
from yolo_sod_b200 import cfg as ycfg, checkpoint

CKPT = os.path.join(os.path.dirname(__file__), "golden", "tiny_ckpt.pt")


def test_reads_reference_checkpoint_without_reference_package():
    had = "ultralytics" in sys.modules
    yaml, sd, names, ckpt = checkpoint.read_checkpoint(CKPT)
    assert had or "ultralytics" not in sys.modules, "loading must not import the reference"
    assert names == {0: "pedestrian", 1: "car", 2: "bicycle"} and yaml["nc"] == 3
    assert ckpt["train_args"]["imgsz"] == 640 and ckpt["epoch"] == 3
    spec = ycfg.get_spec(dict(yaml))
    shapes = ycfg.param_shapes(spec)
    assert list(shapes) == list(sd), "state_dict names / order differ from the graph's parameter inventory"
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in shapes)
    assert all(v.dtype == torch.float32 for k, v in sd.items() if not k.endswith("num_batches_tracked"))   # .float(), tasks.py:945
    assert ycfg.strides_of(spec) == [4, 8]


def test_refuses_foreign_globals():
    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))
    buf = io.BytesIO()
    torch.save({"model": Evil()}, buf)
    path = os.path.join(os.environ.get("TMPDIR", "/tmp"), "ysod_evil.pt")
    open(path, "wb").write(buf.getvalue())
    with pytest.raises(pickle.UnpicklingError):
        checkpoint.torch_safe_load(path)
    os.remove(path)

    class EvilTorch:                                   # a torch.* callable is not a tensor rebuild helper either
        def __reduce__(self):
            return (torch.set_num_threads, (1,))
    buf = io.BytesIO()
    torch.save({"model": EvilTorch()}, buf)
    open(path, "wb").write(buf.getvalue())
    with pytest.raises(pickle.UnpicklingError):
        checkpoint.torch_safe_load(path)
    os.remove(path)


def test_refuses_nested_unrestricted_load():
    """torch.storage._load_from_bytes == torch.load(BytesIO(b), weights_only=False) with the default pickle: a checkpoint that
    REDUCEs it over an embedded payload would run any callable inside the restricted unpickler. It must be refused."""
    import torch.storage

    class Inner:
        def __reduce__(self):
            return (os.getcwd, ())
    inner = io.BytesIO()
    torch.save(Inner(), inner)

    class Outer:
        def __reduce__(self):
            return (torch.storage._load_from_bytes, (inner.getvalue(),))
    buf = io.BytesIO()
    torch.save({"model": Outer()}, buf)
    path = os.path.join(os.environ.get("TMPDIR", "/tmp"), "ysod_nested.pt")
    open(path, "wb").write(buf.getvalue())
    try:
        with pytest.raises(pickle.UnpicklingError):
            checkpoint.torch_safe_load(path)
    finally:
        os.remove(path)
    # prefix / suffix look-alikes are refused too (exact-name allowlist)
    for mod, name in [("torch._utils", "_rebuild_anything"), ("torch", "EvilStorage"), ("torch.storage", "_load_from_bytes")]:
        assert not checkpoint._torch_global_ok(mod, name)


def test_state_dict_equals_live_reference_load():
    from oracle import refshim
    if not refshim.available():
        pytest.skip("live reference not present")
    refshim.load()
    ref = torch.load(CKPT, map_location="cpu", weights_only=False)          # real ultralytics classes
    want = ref["ema"].float().state_dict()
    _, sd, _, _ = checkpoint.read_checkpoint(CKPT)
    assert list(want) == list(sd)
    assert all(torch.equal(want[k], sd[k]) for k in want)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_compiled_checkpoint_matches_oracle_forward(dtype, tol):
    from oracle import model_ref
    from yolo_sod_b200 import synth
    from yolo_sod_b200.model import YOLO
    yolo = YOLO(CKPT, dtype=dtype)
    assert yolo.names[1] == "car" and yolo.model.pt_path == CKPT
    yaml, sd, _, _ = checkpoint.read_checkpoint(CKPT)
    spec = ycfg.get_spec(dict(yaml))
    x = synth.synth_images(2, 128, seed=8)
    y, raw = yolo.model(x.cuda())
    torch.cuda.synchronize()
    y_ref, raw_ref = model_ref.forward(spec, sd, x, ycfg.strides_of(spec))
    for a, b in zip(raw, raw_ref):
        err = (a.float().cpu() - b).abs()
        if tol < 1e-2:
            assert bool((err <= tol * b.abs() + tol * b.abs().max()).all()), float(err.max())
        else:
            assert float(err.norm() / b.norm()) <= tol, float(err.norm() / b.norm())
