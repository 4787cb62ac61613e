"""Deterministic synthetic weights ("calibrated random") for benchmarking and parity tests.

No checkpoints ship with the reference and there is no network, so every run uses fixed-seed random weights of the named
architecture. Plain default init is degenerate for this network (activations collapse to std 4e-7 by
layer 8, zero NMS candidates), so the recipe is: fan-in-normalised Gaussian weights, BN gamma~U(.75,1.25), beta~N(0,.2^2),
and BN running statistics set *per layer* to the scalar mean / variance that layer's conv output has under these weights
(data/synth_calib.json, produced once by tests/golden/make_synth_calib.py), plus a per-level class-bias shift so that a few
percent of the anchors clear conf 0.25. The same state_dict (reference parameter names) feeds the live reference, the
oracle and the CUDA path.
"""
import json
import os
import zlib

import torch

from . import cfg as _cfg

_CALIB = None


def _calib():
    global _CALIB
    if _CALIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "synth_calib.json")
        _CALIB = json.load(open(path)) if os.path.exists(path) else {}
    return _CALIB


def calib_key(cfg_name: str, seed: int) -> str:
    return f"{cfg_name}@seed{seed}"


def synth_state_dict(spec: "_cfg.ModelSpec", cfg_name: str, seed: int = 0, calib: dict = None):
    """Returns an OrderedDict name -> fp32 CPU tensor with the reference's state_dict names.

    `calib`: {bn_prefix: [mean, var], "cls_bias": [per level]} ; defaults to the committed table for (cfg_name, seed).
    """
    if calib is None:
        calib = _calib().get(calib_key(cfg_name, seed), {})
    shapes = _cfg.param_shapes(spec)
    sd = {}
    for name, shape in shapes.items():
        # one generator per tensor, keyed by name: independent of iteration order and of the other tensors
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.long)
        elif leaf == "active_mask":                       # DetectStable: every level active (detect_stable.py:13)
            t = torch.ones(shape, dtype=torch.bool)
        elif ".bn." in name or ".bn1." in name or ".in_proj.1." in name or ".out_proj.1." in name:   # BatchNorm2d (incl. Conv1x1BN's [1])
            pfx = name.rsplit(".", 1)[0]
            m, v = calib.get(pfx, [0.0, 0.4])
            if leaf == "weight":
                t = torch.empty(shape).uniform_(0.75, 1.25, generator=g)
            elif leaf == "bias":
                t = torch.randn(shape, generator=g) * 0.2
            elif leaf == "running_mean":
                t = torch.full(shape, float(m))
            else:
                t = torch.full(shape, float(v))
        elif name.endswith("dfl.conv.weight"):
            t = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)
        elif "norm" in name and leaf == "weight":  # LayerNorm gamma
            t = torch.empty(shape).uniform_(0.75, 1.25, generator=g)
        elif leaf == "gamma":
            t = torch.full(shape, 0.01)
        elif leaf in ("bias", "in_proj_bias"):
            t = torch.randn(shape, generator=g) * 0.1
            if ".cv2." in name and name.endswith(".2.bias"):
                t = torch.ones(shape)                                  # head.py:138 box bias
            if ".cv3." in name and name.endswith(".2.bias"):
                lvl = int(name.split(".cv3.")[1].split(".")[0])
                cb = calib.get("cls_bias")
                t = torch.full(shape, float(cb[lvl]) if cb else -3.0)
        else:  # conv / linear weights: N(0, gain^2 / fan_in)
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            gain = 1.0
            if "in_proj_weight" in name:
                gain = 2.0   # make the attention maps non-uniform
            elif ".fc" in name or "conv_h" in name or "conv_w" in name or "spatial_attention" in name:
                gain = 2.0   # gates away from 0.5
            t = torch.randn(shape, generator=g) * (gain / fan_in ** 0.5)
        sd[name] = t.contiguous()
    return sd


def synth_images(batch: int, imgsz, seed: int = 0):
    """Synthetic input batch: U(0,1) fp32 NCHW, what the reference predictor feeds the model for tensor sources
    (engine/predictor.py:116-134: no /255, no letterbox for tensors)."""
    h, w = (imgsz, imgsz) if isinstance(imgsz, int) else imgsz
    g = torch.Generator().manual_seed(1234 + seed)
    return torch.rand((batch, 3, h, w), generator=g)
