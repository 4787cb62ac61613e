"""Model graph description for the detector families on the hot path, and the YAML -> layer-list resolution.

Mirrors the *behaviour* of the reference's `parse_model` (ultralytics/nn/tasks.py:967-1169) for exactly the module set the
named configs instantiate: Conv, C2f, C3k2, A2C2f, SPPF, SE_Block, CBAM_Block, CA_Block, SwinBlock,
A2_Attn, MambaBlock (GLU fallback), nn.Upsample, Concat, Detect / DetectStable. Anything else raises -- there is no silent fallback.

A user can pass their own reference YAML (path or dict); the two architectures the benchmark names are also built in
(as plain Python data, so nothing from the reference tree is needed at run time):

    "yolov12-sod-fusion-v5-simple"      ultralytics/cfg/models/new/yolov12-sod-fusion-v5-simple.yaml   (13.56 M params)
    "yolov12{n,s,m}"                    ultralytics/cfg/models/v12/yolov12.yaml + scale
    "E1" .. "E6"                        ultralytics/cfg/models/new/E1..E6.yaml (the ablation ladder, derived from the full model)
    "yolov12-sod-fusion-v5-stable"      ultralytics/cfg/models/new/yolov12-sod-fusion-v5-stable.yaml (DetectStable head, 14.21 M params)
    "yolov12-sod-fusion-v5"             ultralytics/cfg/models/new/yolov12-sod-fusion-v5.yaml (-simple + MambaBlock with its GLU fallback, 14.04 M)
"""
import math
import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Union

UP = ["nn.Upsample", [None, 2, "nearest"]]

# [from, repeats, module, args] rows, same meaning as the reference YAML rows
_SOD_SIMPLE = {
    "nc": 10, "depth_multiple": 0.33, "width_multiple": 0.50, "ch": 3,
    "backbone": [
        [-1, 1, "Conv", [64, 3, 2]], [-1, 1, "SE_Block", [64]], [-1, 1, "Conv", [128, 3, 2]], [-1, 3, "C2f", [128, True]],
        [-1, 1, "CBAM_Block", [128, 16]], [-1, 1, "Conv", [256, 3, 2]], [-1, 6, "C2f", [256, True]], [-1, 1, "Conv", [512, 3, 2]],
        [-1, 3, "C2f", [512, True]], [-1, 1, "SwinBlock", [4, 7]], [-1, 1, "Conv", [1024, 3, 2]], [-1, 2, "C2f", [1024, True]],
        [-1, 1, "A2_Attn", [8, 8]], [-1, 1, "SPPF", [1024, 5]],
    ],
    "neck": [
        [-1, 1, "Conv", [512, 1, 1]], [-1, 1] + UP, [[-1, 9], 1, "Concat", [1]], [-1, 3, "C2f", [512, True]],
        [-1, 1, "CBAM_Block", [512, 16]], [-1, 1, "Conv", [256, 1, 1]], [-1, 1] + UP, [[-1, 6], 1, "Concat", [1]],
        [-1, 3, "C2f", [256, True]], [-1, 1, "SE_Block", [256]], [-1, 1, "Conv", [128, 1, 1]], [-1, 1] + UP,
        [[-1, 3], 1, "Concat", [1]], [-1, 3, "C2f", [128, True]], [-1, 1, "SwinBlock", [2, 7]],
        [-1, 1, "Conv", [256, 3, 2]], [[-1, 23], 1, "Concat", [1]], [-1, 3, "C2f", [256, True]], [-1, 1, "CA_Block", [256]],
        [-1, 1, "Conv", [512, 3, 2]], [[-1, 18], 1, "Concat", [1]], [-1, 3, "C2f", [512, True]],
        [-1, 1, "Conv", [1024, 3, 2]], [[-1, 13], 1, "Concat", [1]], [-1, 2, "C2f", [1024, True]],
    ],
    "head": [[[28, 32, 35, 38], 1, "Detect", ["nc"]]],
}

_YOLOV12 = {
    "nc": 80,
    "scales": {"n": [0.50, 0.25, 1024], "s": [0.50, 0.50, 1024], "m": [0.50, 1.00, 512], "l": [1.00, 1.00, 512],
               "x": [1.00, 1.50, 512]},
    "backbone": [
        [-1, 1, "Conv", [64, 3, 2]], [-1, 1, "Conv", [128, 3, 2, 1, 2]], [-1, 2, "C3k2", [256, False, 0.25]],
        [-1, 1, "Conv", [256, 3, 2, 1, 4]], [-1, 2, "C3k2", [512, False, 0.25]], [-1, 1, "Conv", [512, 3, 2]],
        [-1, 4, "A2C2f", [512, True, 4]], [-1, 1, "Conv", [1024, 3, 2]], [-1, 4, "A2C2f", [1024, True, 1]],
    ],
    "head": [
        [-1, 1] + UP, [[-1, 6], 1, "Concat", [1]], [-1, 2, "A2C2f", [512, False, -1]],
        [-1, 1] + UP, [[-1, 4], 1, "Concat", [1]], [-1, 2, "A2C2f", [256, False, -1]],
        [-1, 1, "Conv", [256, 3, 2]], [[-1, 11], 1, "Concat", [1]], [-1, 2, "A2C2f", [512, False, -1]],
        [-1, 1, "Conv", [512, 3, 2]], [[-1, 8], 1, "Concat", [1]], [-1, 2, "C3k2", [1024, True]],
        [[14, 17, 20], 1, "Detect", ["nc"]],
    ],
}



def _ablate(base: dict, drop, detect_from=None) -> dict:
    """An ablation of the SOD architecture: `base` without the layers whose indices are in `drop`. References (`from`) are
    re-indexed; a reference to a dropped layer falls through to the nearest kept layer before it -- which is how the
    reference's hand-written ablation YAMLs (ultralytics/cfg/models/new/E1..E6.yaml) route around the removed attention blocks."""
    rows = base["backbone"] + base.get("neck", []) + base["head"]
    drop = set(drop)
    new_index, k = {}, 0
    for i in range(len(rows)):
        if i not in drop:
            new_index[i] = k
            k += 1

    def remap(i):
        while i in drop:
            i -= 1
        return new_index[i]

    out = []
    for i, (f, n, m, args) in enumerate(rows):
        if i in drop:
            continue
        if m == "Detect" and detect_from is not None:
            f = [remap(j) for j in detect_from]
        elif isinstance(f, list):
            f = [(-1 if j == -1 else remap(j)) for j in f]
        elif f != -1:
            f = remap(f)
        out.append([f, n, m, list(args)])
    nb = sum(1 for i in range(len(base["backbone"])) if i not in drop)
    return {"nc": base["nc"], "depth_multiple": base["depth_multiple"], "width_multiple": base["width_multiple"], "ch": base["ch"],
            "backbone": out[:nb], "neck": out[nb:-1], "head": out[-1:]}


# attention blocks of the full model by layer index: SE {1, 23}, CBAM {4, 18}, Swin {9, 28}, A2 {12}, CA {32}
_SE, _CBAM, _SWIN, _A2, _CA = {1, 23}, {4, 18}, {9, 28}, {12}, {32}
_P2_PATH = {24, 25, 26, 27, 29, 30, 31}   # top-down P3->P2 branch and the P2->P3 bottom-up step

# ultralytics/cfg/models/new/yolov12-sod-fusion-v5-stable.yaml: the "stable" training variant -- no SE/CBAM/CA/A2 blocks, Swin blocks
# at P5 (8 heads) and P2 (4 heads), DetectStable head; its `aux_head:` (DETRAuxHead) section is not part of the inference graph
# (parse_model only walks backbone + neck + head)
_SOD_STABLE = {
    "nc": 10, "depth_multiple": 0.33, "width_multiple": 0.50, "ch": 3,
    "backbone": [
        [-1, 1, "Conv", [64, 3, 2]], [-1, 1, "Conv", [128, 3, 2]], [-1, 3, "C2f", [128, True]], [-1, 1, "Conv", [256, 3, 2]],
        [-1, 6, "C2f", [256, True]], [-1, 1, "Conv", [512, 3, 2]], [-1, 3, "C2f", [512, True]], [-1, 1, "C2f", [512, True]],
        [-1, 1, "Conv", [1024, 3, 2]], [-1, 2, "C2f", [1024, True]], [-1, 1, "SwinBlock", [8, 7]], [-1, 1, "SPPF", [1024, 5]],
    ],
    "neck": [
        [-1, 1, "Conv", [512, 1, 1]], [-1, 1] + UP, [[-1, 6], 1, "Concat", [1]], [-1, 3, "C2f", [512, True]],
        [-1, 1, "Conv", [256, 1, 1]], [-1, 1] + UP, [[-1, 4], 1, "Concat", [1]], [-1, 3, "C2f", [256, True]],
        [-1, 1, "Conv", [128, 1, 1]], [-1, 1] + UP, [[-1, 2], 1, "Concat", [1]], [-1, 3, "C2f", [128, True]],
        [-1, 1, "SwinBlock", [4, 7]],
        [-1, 1, "Conv", [256, 3, 2]], [[-1, 19], 1, "Concat", [1]], [-1, 3, "C2f", [256, True]],
        [-1, 1, "Conv", [512, 3, 2]], [[-1, 15], 1, "Concat", [1]], [-1, 3, "C2f", [512, True]],
        [-1, 1, "Conv", [1024, 3, 2]], [[-1, 11], 1, "Concat", [1]], [-1, 2, "C2f", [1024, True]],
    ],
    "head": [[[24, 27, 30, 33], 1, "DetectStable", ["nc"]]],
}

def _insert_layer(base: dict, at: int, row: list) -> dict:
    """`base` with one extra row inserted before layer index `at` (absolute `from` references at or after it shift by one)."""
    rows = base["backbone"] + base.get("neck", []) + base["head"]
    nb = len(base["backbone"])

    def sh(f):
        return f + 1 if (f != -1 and f >= at) else f

    out = []
    for i, (f, n, m, args) in enumerate(rows):
        if i == at:
            out.append(list(row))
        out.append([[sh(j) for j in f] if isinstance(f, list) else sh(f), n, m, list(args)])
    nb2 = nb + (1 if at <= nb else 0)
    return dict(base, backbone=out[:nb2], neck=out[nb2:-1], head=out[-1:])


# ultralytics/cfg/models/new/yolov12-sod-fusion-v5.yaml = the -simple graph plus a MambaBlock(256, 2) after the P3 C2f (layer 7)
_SOD_V5 = _insert_layer(_SOD_SIMPLE, 7, [-1, 1, "MambaBlock", [256, 2]])

BUILTIN = {
    "yolov12-sod-fusion-v5": _SOD_V5,
    "yolov12-sod-fusion-v5-simple": _SOD_SIMPLE, "yolov12": _YOLOV12, "yolov12-sod-fusion-v5-stable": _SOD_STABLE,
    # the paper's ablation ladder (README.md:131-137; cfg/models/new/E1..E6.yaml): E1 plain PANet with P3-P5 heads, E2 + P2 head,
    # E3 + SE, E4 + CBAM, E5 + Swin, E6 + A2 (the complete model adds CoordAtt)
    "E1": _ablate(_SOD_SIMPLE, _SE | _CBAM | _SWIN | _A2 | _CA | _P2_PATH, detect_from=[22, 35, 38]),
    "E2": _ablate(_SOD_SIMPLE, _SE | _CBAM | _SWIN | _A2 | _CA),
    "E3": _ablate(_SOD_SIMPLE, _CBAM | _SWIN | _A2 | _CA),
    "E4": _ablate(_SOD_SIMPLE, _SWIN | _A2 | _CA),
    "E5": _ablate(_SOD_SIMPLE, _A2 | _CA),
    "E6": _ablate(_SOD_SIMPLE, _CA),
}
ALIASES = {"sod": "yolov12-sod-fusion-v5-simple", "yolov12-sod": "yolov12-sod-fusion-v5-simple"}


def make_divisible(x, divisor):
    """ultralytics/utils/ops.py:130-145."""
    return math.ceil(x / divisor) * divisor


def guess_model_scale(name: str) -> str:
    """tasks.py:1188-1203: yolov12n.yaml -> 'n'."""
    m = re.search(r"yolo[v]?\d+([nslmx])", name)
    return m.group(1) if m else ""


def load_cfg(cfg: Union[str, dict]) -> dict:
    """Resolve a config name / YAML path / dict into a model dict (with `scale` filled in when the name carries one)."""
    import copy
    if isinstance(cfg, dict):
        return copy.deepcopy(cfg)
    name = str(cfg)
    stem = name.rsplit("/", 1)[-1]
    stem = stem[:-5] if stem.endswith(".yaml") else stem
    stem = ALIASES.get(stem, stem)
    if stem in BUILTIN:
        return copy.deepcopy(BUILTIN[stem])
    unified = re.sub(r"(\d+)([nslmx])(.+)?$", r"\1\3", stem)  # yolov12n -> yolov12 (tasks.py:1180)
    if unified in BUILTIN:
        d = copy.deepcopy(BUILTIN[unified])
        d["scale"] = guess_model_scale(stem)
        return d
    if name.endswith(".yaml"):
        import os
        import yaml
        path = name
        if not os.path.exists(path):
            alt = re.sub(r"(\d+)([nslmx])(.+)?$", r"\1\3", path[:-5]) + ".yaml"
            if os.path.exists(alt):
                path = alt
            else:
                raise FileNotFoundError(name)
        with open(path, errors="ignore", encoding="utf-8") as f:
            d = yaml.safe_load(f)
        d["scale"] = d.get("scale") or guess_model_scale(stem)
        return d
    raise ValueError(f"unknown model config {cfg!r}")


@dataclass
class Layer:
    i: int
    f: Union[int, List[int]]
    type: str
    c1: Union[int, List[int]]
    c2: int
    p: Dict[str, Any] = field(default_factory=dict)


@dataclass
class ModelSpec:
    layers: List[Layer]
    save: List[int]
    nc: int
    ch: int
    legacy: bool      # Detect.legacy class attribute as parse_model leaves it (tasks.py:1140-1141)
    scale: str
    yaml: dict


_CONV_LIKE = {"Conv", "C2f", "C3k2", "A2C2f", "SPPF"}
_REPEAT_ARG = {"C2f", "C3k2", "A2C2f"}
_SE_FAMILY = {"SE_Block", "SE", "SwinBlock", "CA_Block", "A2_Attn", "CBAM_Block", "MambaBlock"}


def parse_model(d: dict, ch: int = None, nc: int = None) -> ModelSpec:
    d = dict(d)
    if nc is not None:
        d["nc"] = nc
    nc = d.get("nc")
    ch_in = ch if ch is not None else d.get("ch", 3)
    legacy = True
    max_channels = float("inf")
    depth, width = d.get("depth_multiple", 1.0), d.get("width_multiple", 1.0)
    scales = d.get("scales")
    scale = d.get("scale") or ""
    if scales:
        if not scale:
            scale = tuple(scales.keys())[0]
        depth, width, max_channels = scales[scale]
    if d.get("activation"):
        raise NotImplementedError("custom `activation:` is outside the hot path (SiLU only)")
    chs = [ch_in]
    layers: List[Layer] = []
    save: List[int] = []
    rows = list(d.get("backbone", [])) + list(d.get("neck", [])) + list(d.get("head", []))
    c2 = ch_in
    for i, (f, n, m, args) in enumerate(rows):
        args = [nc if a == "nc" else a for a in args]
        n = max(round(n * depth), 1) if n > 1 else n
        p: Dict[str, Any] = {}
        if m in _CONV_LIKE:
            c1, c2 = chs[f], args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_channels) * width, 8)
            rest = list(args[1:])
            if m in _REPEAT_ARG:
                rest.insert(0, n)
                n = 1
            if m == "Conv":
                names = ["k", "s", "p", "g", "d", "act"]
                p = dict(k=1, s=1, p=None, g=1, d=1, act=True)
                p.update(dict(zip(names, rest)))
            elif m == "C2f":
                p = dict(n=1, shortcut=False, g=1, e=0.5)
                p.update(dict(zip(["n", "shortcut", "g", "e"], rest)))
            elif m == "C3k2":
                legacy = False
                p = dict(n=1, c3k=False, e=0.5, g=1, shortcut=True)
                p.update(dict(zip(["n", "c3k", "e", "g", "shortcut"], rest)))
                if scale in "mlx" and scale:
                    p["c3k"] = True
            elif m == "A2C2f":
                legacy = False
                p = dict(n=1, a2=True, area=1, residual=False, mlp_ratio=2.0, e=0.5, g=1, shortcut=True)
                if scale in "lx" and scale:
                    rest = rest + [True, 1.5]
                p.update(dict(zip(["n", "a2", "area", "residual", "mlp_ratio", "e", "g", "shortcut"], rest)))
            elif m == "SPPF":
                p = dict(k=5)
                p.update(dict(zip(["k"], rest)))
        elif m == "Concat":
            c1 = [chs[x] for x in f]
            c2 = sum(c1)
            p = dict(dim=args[0] if args else 1)
        elif m in _SE_FAMILY:
            c1 = c2 = chs[f]
            if m in ("SE_Block", "SE"):
                p = dict(reduction=args[0] if args else 16)  # YAML arg is the *reduction* (tasks.py:1122-1133)
            elif m == "SwinBlock":
                p = dict(num_heads=args[0] if len(args) > 0 else 4, window_size=args[1] if len(args) > 1 else 7)
            elif m == "CBAM_Block":
                p = dict(reduction=args[1] if len(args) > 1 else 16)   # CBAM_Block(c1, c2=args[0], reduction=args[1])
            elif m == "CA_Block":
                p = dict(reduction=args[1] if len(args) > 1 else 32)   # CA_Block(c1, c2=args[0], reduction=32)
            elif m == "A2_Attn":
                p = dict(num_areas=args[0] if len(args) > 0 else 4, num_heads=args[1] if len(args) > 1 else 4)
            elif m == "MambaBlock":
                # MambaBlock(c, c_hidden=args[0], seq_reduction=args[1]) (tasks.py:1126-1127, blocks_mamba.py:107); compiled with the GLU
                # fallback the reference itself uses whenever mamba_ssm is not importable (blocks_mamba.py:116-165)
                p = dict(c_hidden=args[0] if len(args) > 0 else 256, reduction=args[1] if len(args) > 1 else 2)
        elif m == "nn.Upsample":
            c1 = c2 = chs[f]
            p = dict(scale=args[1], mode=args[2])
            if p["mode"] != "nearest" or args[0] is not None:
                raise NotImplementedError("only nn.Upsample(None, s, 'nearest') is on the hot path")
        elif m in ("Detect", "DetectStable"):
            # DetectStable (nn/modules/detect_stable.py:7-34) is Detect plus a training-only per-level `active_mask` buffer: in eval
            # mode its forward is Detect's, so it maps to the same layer type (the buffer is kept for state_dict compatibility)
            fl = f if isinstance(f, (list, tuple)) else [f]
            c1 = [chs[x] for x in fl]
            c2 = chs[fl[-1]]
            p = dict(nc=args[0], legacy=legacy, stable=(m == "DetectStable"))
            m = "Detect"
        else:
            raise NotImplementedError(f"module {m!r} (layer {i}) is outside the hot path this library covers")
        if n != 1:
            raise NotImplementedError(f"repeat count {n} for module {m!r} (layer {i}) is not used by the supported configs")
        layers.append(Layer(i=i, f=f, type=m, c1=c1, c2=c2, p=p))
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        if i == 0:
            chs = []
        chs.append(c2)
    return ModelSpec(layers=layers, save=sorted(save), nc=nc, ch=ch_in, legacy=legacy, scale=scale, yaml=d)


def get_spec(cfg: Union[str, dict], ch: int = None, nc: int = None) -> ModelSpec:
    return parse_model(load_cfg(cfg), ch=ch, nc=nc)


# ---- parameter inventory: reference state_dict names -> shapes -----------------------------------------------------

def _conv(sh, pfx, c1, c2, k=1, g=1):
    sh[f"{pfx}.conv.weight"] = (c2, c1 // g, k, k)
    for s in ("weight", "bias", "running_mean", "running_var"):
        sh[f"{pfx}.bn.{s}"] = (c2,)
    sh[f"{pfx}.bn.num_batches_tracked"] = ()


def _bottleneck(sh, pfx, c1, c2, g=1, k=(3, 3), e=0.5):
    c_ = int(c2 * e)
    _conv(sh, f"{pfx}.cv1", c1, c_, k[0])
    _conv(sh, f"{pfx}.cv2", c_, c2, k[1], g)


def _c3k(sh, pfx, c1, c2, n=2, g=1, e=0.5, k=3):
    c_ = int(c2 * e)
    _conv(sh, f"{pfx}.cv1", c1, c_, 1)
    _conv(sh, f"{pfx}.cv2", c1, c_, 1)
    _conv(sh, f"{pfx}.cv3", 2 * c_, c2, 1)
    for j in range(n):
        _bottleneck(sh, f"{pfx}.m.{j}", c_, c_, g, (k, k), 1.0)


def _ablock(sh, pfx, dim, mlp_ratio):
    _conv(sh, f"{pfx}.attn.qk", dim, 2 * dim, 1)
    _conv(sh, f"{pfx}.attn.v", dim, dim, 1)
    _conv(sh, f"{pfx}.attn.proj", dim, dim, 1)
    _conv(sh, f"{pfx}.attn.pe", dim, dim, 5, dim)
    hid = int(dim * mlp_ratio)
    _conv(sh, f"{pfx}.mlp.0", dim, hid, 1)
    _conv(sh, f"{pfx}.mlp.1", hid, dim, 1)


def _mha(sh, pfx, e):
    sh[f"{pfx}.in_proj_weight"] = (3 * e, e)
    sh[f"{pfx}.in_proj_bias"] = (3 * e,)
    sh[f"{pfx}.out_proj.weight"] = (e, e)
    sh[f"{pfx}.out_proj.bias"] = (e,)


def detect_channels(nc, ch):
    """head.py:40: c2, c3."""
    return max((16, ch[0] // 4, 64)), max(ch[0], min(nc, 100))


def param_shapes(spec: ModelSpec) -> "Dict[str, tuple]":
    """Every tensor of the reference model's state_dict() (after its first forward, i.e. including the lazily built SE
    weights, smallobj_modules.py:72-82), in the reference's registration order."""
    from collections import OrderedDict
    sh: Dict[str, tuple] = OrderedDict()
    for L in spec.layers:
        P = f"model.{L.i}"
        t, p = L.type, L.p
        if t == "Conv":
            _conv(sh, P, L.c1, L.c2, p["k"], p["g"])
        elif t == "C2f":
            c = int(L.c2 * p["e"])
            _conv(sh, f"{P}.cv1", L.c1, 2 * c, 1)
            _conv(sh, f"{P}.cv2", (2 + p["n"]) * c, L.c2, 1)
            for j in range(p["n"]):
                _bottleneck(sh, f"{P}.m.{j}", c, c, p["g"], (3, 3), 1.0)
        elif t == "C3k2":
            c = int(L.c2 * p["e"])
            _conv(sh, f"{P}.cv1", L.c1, 2 * c, 1)
            _conv(sh, f"{P}.cv2", (2 + p["n"]) * c, L.c2, 1)
            for j in range(p["n"]):
                if p["c3k"]:
                    _c3k(sh, f"{P}.m.{j}", c, c, 2, p["g"])
                else:
                    _bottleneck(sh, f"{P}.m.{j}", c, c, p["g"])
        elif t == "A2C2f":
            c_ = int(L.c2 * p["e"])
            _conv(sh, f"{P}.cv1", L.c1, c_, 1)
            _conv(sh, f"{P}.cv2", (1 + p["n"]) * c_, L.c2, 1)
            if p["a2"] and p["residual"]:
                sh[f"{P}.gamma"] = (L.c2,)
            for j in range(p["n"]):
                if p["a2"]:
                    for q in range(2):
                        _ablock(sh, f"{P}.m.{j}.{q}", c_, p["mlp_ratio"])
                else:
                    _c3k(sh, f"{P}.m.{j}", c_, c_, 2, p["g"])
        elif t == "SPPF":
            _conv(sh, f"{P}.cv1", L.c1, L.c1 // 2, 1)
            _conv(sh, f"{P}.cv2", (L.c1 // 2) * 4, L.c2, 1)
        elif t in ("SE_Block", "SE"):
            hid = max(L.c1 // p["reduction"], 4)
            sh[f"{P}.fc1.weight"] = (hid, L.c1, 1, 1)
            sh[f"{P}.fc1.bias"] = (hid,)
            sh[f"{P}.fc2.weight"] = (L.c1, hid, 1, 1)
            sh[f"{P}.fc2.bias"] = (L.c1,)
        elif t == "CBAM_Block":
            hid = L.c1 // p["reduction"]
            sh[f"{P}.channel_attention.fc.0.weight"] = (hid, L.c1, 1, 1)
            sh[f"{P}.channel_attention.fc.2.weight"] = (L.c1, hid, 1, 1)
            sh[f"{P}.spatial_attention.conv1.weight"] = (1, 2, 7, 7)
        elif t == "CA_Block":
            mip = max(8, L.c1 // p["reduction"])
            sh[f"{P}.conv1.weight"] = (mip, L.c1, 1, 1)
            sh[f"{P}.conv1.bias"] = (mip,)
            for s in ("weight", "bias", "running_mean", "running_var"):
                sh[f"{P}.bn1.{s}"] = (mip,)
            sh[f"{P}.bn1.num_batches_tracked"] = ()
            for s in ("conv_h", "conv_w"):
                sh[f"{P}.{s}.weight"] = (L.c1, mip, 1, 1)
                sh[f"{P}.{s}.bias"] = (L.c1,)
        elif t == "SwinBlock":
            c = L.c1
            sh[f"{P}.dw.weight"] = (c, 1, 3, 3)
            for s in ("weight", "bias"):
                sh[f"{P}.window_attn.norm1.{s}"] = (c,)
            _mha(sh, f"{P}.window_attn.attn", c)
            for s in ("weight", "bias"):
                sh[f"{P}.window_attn.norm2.{s}"] = (c,)
            sh[f"{P}.window_attn.mlp.0.weight"] = (2 * c, c)
            sh[f"{P}.window_attn.mlp.0.bias"] = (2 * c,)
            sh[f"{P}.window_attn.mlp.2.weight"] = (c, 2 * c)
            sh[f"{P}.window_attn.mlp.2.bias"] = (c,)
            sh[f"{P}.pw.weight"] = (c, c, 1, 1)
            for s in ("weight", "bias", "running_mean", "running_var"):
                sh[f"{P}.bn.{s}"] = (c,)
            sh[f"{P}.bn.num_batches_tracked"] = ()
        elif t == "MambaBlock":
            c, ch = L.c1, p["c_hidden"]
            hid = 2 * ch                                   # GLUBlock(c_hidden, expansion=2)
            for pre, ci, co in ((f"{P}.in_proj", c, ch), (f"{P}.out_proj", ch, c)):      # Conv1x1BN = Sequential(conv, bn, act)
                sh[f"{pre}.0.weight"] = (co, ci, 1, 1)
                for s_ in ("weight", "bias", "running_mean", "running_var"):
                    sh[f"{pre}.1.{s_}"] = (co,)
                sh[f"{pre}.1.num_batches_tracked"] = ()
            sh[f"{P}.fallback.pw1.weight"] = (2 * hid, ch, 1, 1)
            sh[f"{P}.fallback.dw.weight"] = (hid, 1, 3, 3)
            for s_ in ("weight", "bias", "running_mean", "running_var"):
                sh[f"{P}.fallback.bn.{s_}"] = (hid,)
            sh[f"{P}.fallback.bn.num_batches_tracked"] = ()
            sh[f"{P}.fallback.pw2.weight"] = (ch, hid, 1, 1)
        elif t == "A2_Attn":
            c = L.c1
            _conv(sh, f"{P}.proj", c, c, 1)
            _mha(sh, f"{P}.attention", c)
            _conv(sh, f"{P}.out_proj", c, c, 1)
            for s in ("weight", "bias"):
                sh[f"{P}.layer_norm.{s}"] = (c,)
        elif t == "Detect":
            nc, ch = p["nc"], L.c1
            c2, c3 = detect_channels(nc, ch)
            if p.get("stable"):
                sh[f"{P}.active_mask"] = (len(ch),)
            for i, x in enumerate(ch):
                _conv(sh, f"{P}.cv2.{i}.0", x, c2, 3)
                _conv(sh, f"{P}.cv2.{i}.1", c2, c2, 3)
                sh[f"{P}.cv2.{i}.2.weight"] = (64, c2, 1, 1)
                sh[f"{P}.cv2.{i}.2.bias"] = (64,)
            for i, x in enumerate(ch):
                if p["legacy"]:
                    _conv(sh, f"{P}.cv3.{i}.0", x, c3, 3)
                    _conv(sh, f"{P}.cv3.{i}.1", c3, c3, 3)
                else:
                    _conv(sh, f"{P}.cv3.{i}.0.0", x, x, 3, x)
                    _conv(sh, f"{P}.cv3.{i}.0.1", x, c3, 1)
                    _conv(sh, f"{P}.cv3.{i}.1.0", c3, c3, 3, c3)
                    _conv(sh, f"{P}.cv3.{i}.1.1", c3, c3, 1)
                sh[f"{P}.cv3.{i}.2.weight"] = (nc, c3, 1, 1)
                sh[f"{P}.cv3.{i}.2.bias"] = (nc,)
            sh[f"{P}.dfl.conv.weight"] = (1, 16, 1, 1)
    return sh


def strides_of(spec: ModelSpec) -> List[int]:
    """Detect strides, computed from the graph instead of the reference's 256x256 dry run (tasks.py:366-370)."""
    red = {}
    for L in spec.layers:
        src = L.f if isinstance(L.f, int) else L.f[0]
        r_in = 1 if (L.i == 0) else red[src % L.i if src != -1 else L.i - 1]
        if L.type == "Conv":
            red[L.i] = r_in * L.p["s"]
        elif L.type == "nn.Upsample":
            red[L.i] = r_in / L.p["scale"]
        elif L.type == "Detect":
            return [int(red[x]) for x in L.f]
        else:
            red[L.i] = r_in
    raise ValueError("no Detect layer")
