"""Checkpoint ingestion: compile a trained reference `.pt` straight into the B200 engine.

    attempt_load_one_weight(weight, device, inplace, fuse)   ultralytics/nn/tasks.py:941-964
    torch_safe_load(weight)                                    ultralytics/nn/tasks.py:860-900

A reference checkpoint is a pickled dict whose "model" / "ema" entries are whole pickled `DetectionModel` module trees
(`ultralytics.nn.tasks.DetectionModel` and the `ultralytics.nn.modules.*` classes). This package does not ship those classes and
must not need the reference installed, so the unpickler resolves every `ultralytics.*` global to an inert stand-in that only
receives the pickled `__dict__`; real `torch.nn` leaves (Conv2d, BatchNorm2d, MultiheadAttention ...) unpickle as themselves. From
that tree we read exactly what the engine compiles from: the model YAML dict (`model.yaml`), the reference-named state_dict
(parameters + persistent buffers, walked in module order) and `names` / `nc`. No reference code is executed.

Globals outside {tensor / storage rebuild helpers, torch.nn module classes, plain containers, numpy array reconstruction, a few
builtins, ultralytics.* stand-ins} are refused -- in particular arbitrary torch.* callables.
"""
import pickle
from collections import OrderedDict
from typing import Dict, Tuple

import torch

_SAFE_BUILTINS = {"set", "frozenset", "slice", "range", "complex", "bytearray", "dict", "list", "tuple", "int", "float", "bool", "str"}
# Plain data containers / array reconstruction only. torch is NOT blanket-allowed (torch.hub.*, torch.load ... are callables a
# crafted pickle could REDUCE): tensor / storage / parameter rebuild helpers, dtype-like singletons and nn.Module classes only.
_SAFE_GLOBALS = {
    ("collections", "OrderedDict"), ("collections", "defaultdict"), ("pathlib", "PosixPath"), ("pathlib", "PurePosixPath"),
    ("types", "SimpleNamespace"), ("argparse", "Namespace"), ("_codecs", "encode"), ("copyreg", "_reconstructor"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "scalar"),
    ("numpy._core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "scalar"),
    ("torch", "Size"), ("torch", "device"), ("torch", "Tensor"), ("torch.nn.parameter", "Parameter"),
    ("torch.serialization", "_get_layout"),
}
_TORCH_DTYPES = {"float32", "float16", "bfloat16", "float64", "int64", "int32", "int16", "int8", "uint8", "bool"}


# exact names only (no prefix / suffix matching): the helpers torch.save emits for tensors, parameters and storages
_TORCH_REBUILD = {
    ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor_v3"),
    ("torch._utils", "_rebuild_parameter"), ("torch._utils", "_rebuild_parameter_with_state"),
    ("torch._utils", "_rebuild_qtensor"), ("torch._tensor", "_rebuild_from_type_v2"),
}
_TORCH_STORAGES = {"FloatStorage", "HalfStorage", "BFloat16Storage", "DoubleStorage", "LongStorage", "IntStorage", "ShortStorage",
                   "CharStorage", "ByteStorage", "BoolStorage", "UntypedStorage", "TypedStorage"}


def _torch_global_ok(module: str, name: str) -> bool:
    """torch.storage._load_from_bytes is deliberately NOT allowed: it is `torch.load(BytesIO(b), weights_only=False)` with the
    default, unrestricted pickle, so a checkpoint could REDUCE it over an embedded payload and run any callable. Zip-format
    checkpoints (everything torch >= 1.6 writes) never reference it."""
    if (module, name) in _SAFE_GLOBALS or (module, name) in _TORCH_REBUILD:
        return True
    if module in ("torch", "torch.storage") and name in _TORCH_STORAGES:
        import inspect
        return inspect.isclass(getattr(__import__(module, fromlist=[name]), name, None))
    if module == "torch":
        return name in _TORCH_DTYPES                                              # torch.float32 ...
    if module.startswith("torch.nn.modules."):
        import importlib
        import inspect
        obj = getattr(importlib.import_module(module), name, None)
        return inspect.isclass(obj) and issubclass(obj, torch.nn.Module)          # Conv2d, BatchNorm2d, MultiheadAttention ...
    return False


_stub_cache: Dict[Tuple[str, str], type] = {}


class StubObject:
    """Stand-in for a pickled `ultralytics.*` object: keeps the pickled state, runs no code."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        elif isinstance(state, tuple) and len(state) == 2:      # (dict state, slots state)
            for part in state:
                if isinstance(part, dict):
                    self.__dict__.update(part)

    def __repr__(self):
        return f"<stub {type(self).__module__}.{type(self).__qualname__}>"


def _stub_class(module: str, name: str) -> type:
    key = (module, name)
    if key not in _stub_cache:
        _stub_cache[key] = type(name, (StubObject,), {"__module__": module})
    return _stub_cache[key]


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "ultralytics" or module.startswith("ultralytics.") or module == "__main__":
            return _stub_class(module, name)
        if module in ("builtins", "__builtin__"):
            if name in _SAFE_BUILTINS:
                return super().find_class("builtins", name)
            raise pickle.UnpicklingError(f"refusing builtins.{name} in a checkpoint")
        if (module, name) in _SAFE_GLOBALS or (module.startswith("torch") and _torch_global_ok(module, name)):
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"refusing global {module}.{name} in a checkpoint")


class _PickleModule:
    """What torch.load expects from `pickle_module`."""
    __name__ = "yolo_sod_b200_checkpoint_pickle"
    Unpickler = _Unpickler
    UnpicklingError = pickle.UnpicklingError

    @staticmethod
    def load(f, **kw):
        return _Unpickler(f, **kw).load()


def torch_safe_load(weight: str) -> dict:
    """tasks.py:860-900 without the reference package: returns the checkpoint dict (module entries are stand-in trees)."""
    ckpt = torch.load(weight, map_location="cpu", pickle_module=_PickleModule, weights_only=False)
    if not isinstance(ckpt, dict):
        ckpt = {"model": getattr(ckpt, "model", ckpt)}      # tasks.py:893-898: a bare pickled model
    return ckpt


def module_state_dict(root) -> "OrderedDict[str, torch.Tensor]":
    """`nn.Module.state_dict()` for a tree that mixes stand-ins and real torch modules: own parameters, own persistent buffers,
    then children, depth first in registration order (torch/nn/modules/module.py `_save_to_state_dict` / `state_dict`)."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def walk(m, prefix):
        d = m.__dict__
        for k, v in (d.get("_parameters") or {}).items():
            if v is not None:
                out[prefix + k] = v.detach()
        skip = d.get("_non_persistent_buffers_set") or set()
        for k, v in (d.get("_buffers") or {}).items():
            if v is not None and k not in skip:
                out[prefix + k] = v.detach()
        for k, c in (d.get("_modules") or {}).items():
            if c is not None:
                walk(c, prefix + k + ".")

    walk(root, "")
    return out


def read_checkpoint(weight: str):
    """Returns (yaml dict, state_dict fp32, names, ckpt): `ema` is preferred over `model` as in tasks.py:945."""
    ckpt = torch_safe_load(weight)
    model = ckpt.get("ema") or ckpt["model"]
    yaml = getattr(model, "yaml", None)
    if not isinstance(yaml, dict):
        raise ValueError(f"{weight}: the pickled model carries no `yaml` dict (tasks.py:342); cannot rebuild the graph")
    sd = OrderedDict((k, v.float() if v.is_floating_point() else v) for k, v in module_state_dict(model).items())   # .float(): tasks.py:945
    names = getattr(model, "names", None)
    return yaml, sd, names, ckpt


def attempt_load_one_weight(weight: str, device="cuda:0", inplace=True, fuse=False, dtype=torch.bfloat16, **kw):
    """tasks.py:941-964: `(model, ckpt)`, with `model` a compiled B200 DetectionModel carrying the checkpoint's weights, names,
    yaml and `pt_path`. `fuse` / `inplace` are accepted for signature parity (BN is always folded; there are no in-place modules)."""
    from .model import DetectionModel
    yaml, sd, names, ckpt = read_checkpoint(weight)
    model = DetectionModel(dict(yaml), weights=sd, dtype=dtype, device=device or "cuda:0", **kw)
    if isinstance(names, dict):
        model.names = dict(names)
    elif isinstance(names, (list, tuple)):
        model.names = dict(enumerate(names))
    model.pt_path = weight
    model.args = ckpt.get("train_args", {})
    return model, ckpt
