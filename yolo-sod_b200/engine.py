"""Compiles a detector (layer spec + reference-named state_dict) into a static program of sm_100a kernel launches.

This is the host side of the forward hot path. It replaces the reference's Python layer loop
(ultralytics/nn/tasks.py:165-192 `_predict_once`), `BaseModel.fuse` (tasks.py:227-255 + utils/torch_utils.py:238-265
BN folding, eps 1e-3) and the per-module forward methods, by:
  * folding every BatchNorm (incl. the raw ones in SwinBlock.bn / CA_Block.bn1) into conv weights + fp32 bias,
  * repacking weights K-major ([Cout][kh][kw][Cin]) in bf16 (fp32 in fp32 mode),
  * planning NHWC activation buffers so that producers write straight into channel slices of their consumer's
    buffer (Concat / chunk / cat never copy),
  * emitting one flat list of C-ABI calls (lib.py) that is replayed per batch, optionally under a CUDA graph.
Nothing here computes on the CPU at run time; if libysod.so or the GPU is missing it raises.
"""
import ctypes as C
import math
import os
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import cfg as _cfg
from . import lib as _lib

BN_EPS = 1e-3   # utils/torch_utils.py:416-418
LN_EPS = 1e-5   # torch.nn.LayerNorm default
RAW_CS = 80     # pixel stride of the raw head maps (64 box + nc class channels, padded to a multiple of 8)


@dataclass
class View:
    """NHWC view: element (n,h,w,c) lives at buf[off + ((n*H + h)*W + w)*cs + c]."""
    buf: torch.Tensor
    off: int
    N: int
    H: int
    W: int
    C: int
    cs: int
    sub: int = 1   # > 1: this (N,H,W) view samples every `sub`-th pixel of an (N, H*sub, W*sub) buffer (a conv whose output only
                   # exists nearest-upsampled, see Program._build); for layer_output() only, never handed to a kernel

    def ptr(self):
        return C.c_void_p(self.buf.data_ptr() + self.off * self.buf.element_size())

    def slice(self, c0, c1):
        return View(self.buf, self.off + c0, self.N, self.H, self.W, c1 - c0, self.cs)

    def torch_nhwc(self):
        u = self.sub
        t = self.buf[self.off:].as_strided((self.N, self.H, self.W, self.C),
                                           (self.H * u * self.W * u * self.cs, u * self.W * u * self.cs, u * self.cs, 1))
        return t

    def torch_nchw(self):
        return self.torch_nhwc().permute(0, 3, 1, 2)


@dataclass
class GatedView:
    """Output of an SE block whose `x * gate` was folded into the weights of the conv that consumes it (Program._build): the gated
    map is never materialised; layer_output() reconstructs it from the block's input and the gate for the parity tests."""
    base: View
    gate: torch.Tensor   # (N, C) fp32

    def torch_nchw(self):
        return self.base.torch_nchw().float() * self.gate[:, :, None, None]


class Program:
    """A compiled forward for one (batch, H, W): buffers + launch list."""

    def __init__(self, model: "B200DetectionModel", B: int, H: int, W: int, src_u8: bool = False, want_raw: bool = True):
        self.m = model
        self.B, self.H, self.W = B, H, W
        self.want_raw = want_raw   # False: the predict path -- only `y` is consumed, the fp32 raw maps are not materialised
        self.src_u8 = src_u8   # input is (B,H,W,3) uint8 BGR frames (predictor.py:116-134) instead of (B,3,H,W) float
        self.dev = model.device
        self.dt = model.dtype
        self.half = self.dt == torch.float16      # fp16 build of the library (the reference's half=True mode)
        self.lib = _lib.load(self.half)
        self.code = _lib.BF16 if self.dt in (torch.bfloat16, torch.float16) else _lib.F32   # code 1 = the build's 16-bit storage type
        self.ops: List[tuple] = []       # (cfunc, args-without-stream, entry point name)
        self.sched: List[tuple] = []     # launch schedule: ("op", op index, lane) | ("record", mark, lane) | ("wait", mark, lane)
        self._lane = 0                   # stream lane new ops are issued on (0 = main chain; 1.. = Detect level branches)
        self.n_lanes = 1
        self.keep: List[object] = []     # tensors / handles that must stay alive
        self.tc_handles: List[C.c_void_p] = []
        self.n_launches = 0
        self.n_tc = 0
        self.tc_flops = 0        # algorithmic FLOPs (2*M*N*K, unpadded) of all tcgen05 conv launches of one forward
        self.op_flops: List[float] = []   # per op, parallel to self.ops (0 for non-GEMM ops)
        self.op_desc: List[str] = []      # per op, human-readable shape
        self._ctx = ""
        self.layer_out: Dict[int, View] = {}
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        # The stem reads its image through a device pointer slot (ysod_set_ptr), so a forward consumes the caller's tensor in
        # place, as the reference does; `img` is a private staging buffer, allocated only for inputs that cannot be bound
        # directly (host tensors, other dtypes / layouts).
        self.img_shape = (B, H, W, 3) if src_u8 else (B, 3, H, W)
        self.img_dtype = torch.uint8 if src_u8 else torch.float32
        self.img: Optional[torch.Tensor] = None
        self.img_slot = torch.zeros(1, device=self.dev, dtype=torch.int64)
        self.img_indirect = False
        self._bound = 0
        self._build()
        if not self.img_indirect:
            self.img = torch.zeros(self.img_shape, device=self.dev, dtype=self.img_dtype)
            self._rebind_direct()

    # ---- buffers -----------------------------------------------------------------------------------------
    def new(self, N, H, W, Cc, dtype=None, zero=False):
        dtype = dtype or self.dt
        n = N * H * W * Cc
        buf = (torch.zeros if zero else torch.empty)(n + 64, device=self.dev, dtype=dtype)
        self.keep.append(buf)
        return View(buf, 0, N, H, W, Cc, Cc)

    def f32(self, *shape):
        t = torch.empty(shape, device=self.dev, dtype=torch.float32)
        self.keep.append(t)
        return t

    def dev_t(self, t, dtype=None):
        t = t.detach().to(device=self.dev, dtype=dtype or torch.float32).contiguous()
        self.keep.append(t)
        return t

    def call(self, name, *args):
        _lib.check(getattr(self.lib, name)(*args), name, self.half)

    def emit(self, name, *args, flops=0.0, desc=""):
        self.ops.append((getattr(self.lib, name), args, name))
        self.sched.append(("op", len(self.ops) - 1, self._lane))
        self.op_flops.append(float(flops))
        self.op_desc.append(f"{self._ctx} {desc}".strip())
        self.n_launches += 1

    # ---- conv emission -----------------------------------------------------------------------------------
    def folded(self, pfx):
        sd = self.m.sd
        w = sd[f"{pfx}.conv.weight"].float()
        g, b = sd[f"{pfx}.bn.weight"].float(), sd[f"{pfx}.bn.bias"].float()
        mu, var = sd[f"{pfx}.bn.running_mean"].float(), sd[f"{pfx}.bn.running_var"].float()
        s = g / torch.sqrt(var + BN_EPS)
        return w * s.view(-1, 1, 1, 1), b - mu * s

    def conv(self, x: View, w: torch.Tensor, bias: torch.Tensor, k, s, g, act, out: View, res: View = None,
             out_f32=False, pad=None, up2=False, gate=None, alg_flops=None, no_store=False, no_duo=False):
        """w: (Cout, Cin/g, k, k) fp32 with BN folded; bias fp32 (Cout). up2: `out` is the 2x nearest-upsampled destination
        (tensor-core path only; the caller checks `tc_eligible`). gate: (N, Cin) fp32 device tensor of an SE block folded into this conv
        (per-image weights W * gate[n], rebuilt every forward by ysod_scale_weights; tensor-core path only)."""
        pad = k // 2 if pad is None else pad
        Cout, Cin = w.shape[0], x.C
        assert w.shape[1] * g == Cin, (w.shape, Cin, g)
        Ho, Wo = (x.H + 2 * pad - k) // s + 1, (x.W + 2 * pad - k) // s + 1
        us = 2 if up2 else 1
        assert (out.N, out.H, out.W, out.C) == (x.N, us * Ho, us * Wo, Cout), ((out.N, out.H, out.W, out.C), (x.N, Ho, Wo, Cout))
        actc = _lib.ACT[act]
        odt = _lib.F32 if out_f32 else self.code
        if res is not None:
            assert (res.N, res.H, res.W, res.C) == (out.N, out.H, out.W, out.C)
        if (g > 1 and not (g == Cin and g == Cout) and self.m.use_tc and self.code == _lib.BF16 and (Cin // g) % 32 == 0
                and (Cout // g) % 8 == 0 and res is None):
            # grouped conv (yolov12.yaml:20,22) = g dense convs on channel slices of the same NHWC buffers: tensor-core kernel per group
            ci, co = Cin // g, Cout // g
            for j in range(g):
                self.conv(x.slice(j * ci, (j + 1) * ci), w[j * co:(j + 1) * co], bias[j * co:(j + 1) * co], k, s, 1, act,
                          out.slice(j * co, (j + 1) * co), None, out_f32, pad)
            return
        use_tc = (self.m.use_tc and self.code == _lib.BF16 and g == 1 and k in (1, 3) and s in (1, 2) and pad == k // 2
                  and Cin % 32 == 0 and (s == 1 or (x.H % 2 == 0 and x.W % 2 == 0)))
        assert use_tc or not (up2 or gate is not None or no_store), "fused upsample / folded SE gate / decode-only need the tensor-core conv"
        if (use_tc and k == 1 and s == 1 and not up2 and gate is None and not out_f32 and not no_store and x.sub == 1 and out.sub == 1
                and (x.H * x.W) % 128 != 0 and x.N * x.H * x.W >= 128):
            # A 1x1 conv is a GEMM over pixels and NHWC views are pixel-contiguous across rows and images, so the map is handed to the
            # kernel as ONE row of N*H*W pixels: every M = 128 tile is full (2-D tiles of a 40 x 40 map waste 9 % of their rows, of a
            # 20 x 20 map 17 %) and the tile count -- hence the number of waves over the 148 SMs -- drops by as much.
            npx = x.N * x.H * x.W
            flat = lambda v: View(v.buf, v.off, 1, 1, npx, v.C, v.cs)
            x, out = flat(x), flat(out)
            res = flat(res) if res is not None else None
            Ho, Wo = 1, npx
        if use_tc:
            cpad = (Cout + 15) // 16 * 16
            wk = torch.zeros((cpad, k * k * Cin), dtype=torch.float32)
            wk[:Cout] = w.permute(0, 2, 3, 1).reshape(Cout, -1)
            bk = torch.zeros(cpad, dtype=torch.float32)
            bk[:Cout] = bias
            bd = self.dev_t(bk)
            mode = _lib.CONV_UP2 if up2 else 0
            if not self.m.conv_pair:
                mode |= _lib.CONV_NO_PAIR
            if no_duo or not self.m.conv_duo:
                mode |= _lib.CONV_NO_DUO
            if os.environ.get("YSOD_CONV_DEBUG"):
                mode |= int(os.environ["YSOD_CONV_DEBUG"]) << 8   # A/B switches of the conv kernel (debug bits, tc_conv.cu)
            if os.environ.get("YSOD_TRACE_OP", "") == str(len(self.ops)):
                mode |= 32 << 8   # profiling aid: CTA 0 of this op logs its pipeline events (tools/trace_in_graph.py)
            if no_store:
                mode |= _lib.CONV_NO_STORE | _lib.CONV_NO_SPLIT_STAGING
            if gate is None:
                wd = self.dev_t(wk, self.dt)
            else:
                # one weight matrix per image: fp32 master weights x gate[n] -> bf16, rebuilt by a small launch before the conv
                wm = self.dev_t(wk)
                wd = torch.empty((x.N, cpad, k * k * Cin), device=self.dev, dtype=self.dt)
                self.keep.append(wd)
                self.emit("ysod_scale_weights", _lib.ptr(wm), cpad, k * k * Cin, Cin, _lib.ptr(gate), x.N, _lib.ptr(wd), desc="SE gate -> conv weights")
                mode |= _lib.CONV_IMG_WEIGHTS
            h = C.c_void_p()
            self.call("ysod_conv_tc_create_ex", C.byref(h), x.ptr(), x.N, x.H, x.W, Cin, x.cs, _lib.ptr(wd), _lib.ptr(bd), Cout, cpad,
                      k, s, out.ptr(), odt, out.cs, res.ptr() if res is not None else None, res.cs if res is not None else 0, actc,
                      mode)
            self.tc_handles.append(h)
            self.ops.append((self.lib.ysod_conv_tc_run, (h,), "ysod_conv_tc_run"))
            self.sched.append(("op", len(self.ops) - 1, self._lane))
            fl = 2.0 * x.N * Ho * Wo * Cout * k * k * Cin if alg_flops is None else float(alg_flops)   # algorithmic (structural zeros excluded)
            self.op_flops.append(fl)
            info = (C.c_int * 8)()
            self.call("ysod_conv_tc_info", h, info)
            self.op_desc.append(f"{self._ctx} tc {Cin}->{Cout} k{k}s{s}{'up2' if up2 else ''}{'+SEgate' if gate is not None else ''} @{Ho}x{Wo} N{x.N} tile{info[0]}x{info[1]} BN{info[2]} BK{info[3]} "
                                f"st{info[4]} grid{info[5]}x{info[6]} smem{info[7]}")
            self.tc_flops += fl
            self.n_launches += 1
            self.n_tc += 1
            return
        if g == Cin and g == Cout and g > 1:
            wd = self.dev_t(w.view(Cout, k, k).permute(1, 2, 0), self.dt)  # [k][k][C]
            bd = self.dev_t(bias)
            assert not out_f32
            self.emit("ysod_dwconv", x.ptr(), self.code, x.N, x.H, x.W, Cin, x.cs, _lib.ptr(wd), _lib.ptr(bd), k, s, pad, out.ptr(),
                      out.cs, res.ptr() if res is not None else None, res.cs if res is not None else 0, actc,
                      flops=2.0 * x.N * Ho * Wo * Cout * k * k, desc=f"dw {Cin} k{k} @{Ho}x{Wo}")
            return
        wd = self.dev_t(w.permute(0, 2, 3, 1), self.dt)  # [Cout][k][k][Cin/g]
        bd = self.dev_t(bias)
        self.emit("ysod_conv_direct", x.ptr(), self.code, x.N, x.H, x.W, Cin, x.cs, _lib.ptr(wd), _lib.ptr(bd), Cout, k, s, pad, g,
                  out.ptr(), odt, out.cs, res.ptr() if res is not None else None, res.cs if res is not None else 0, actc,
                  flops=2.0 * x.N * Ho * Wo * Cout * k * k * Cin / g, desc=f"direct {Cin}->{Cout} k{k}s{s}g{g} @{Ho}x{Wo}")

    def conv_bn(self, x, pfx, k=1, s=1, g=1, act=True, out=None, res=None, pad=None, up2=False, gate=None, no_duo=False):
        """Reference `Conv` wrapper (conv.py:37-55) with BN folded."""
        w, b = self.folded(pfx)
        if out is None:
            p = k // 2 if pad is None else pad
            out = self.new(x.N, (x.H + 2 * p - k) // s + 1, (x.W + 2 * p - k) // s + 1, w.shape[0])
        self.conv(x, w, b, k, s, g, "silu" if act else "none", out, res, pad=pad, up2=up2, gate=gate, no_duo=no_duo)
        return out

    def tc_eligible(self, L, cin, h, w):
        """True when the top-level Conv layer L runs on the tensor-core kernel (dense, k in {1,3}, default padding): the precondition
        for writing a fused nn.Upsample(2) output or taking per-image (SE-gated) weights."""
        p = L.p
        return (self.m.use_tc and self.code == _lib.BF16 and p["g"] == 1 and p["k"] in (1, 3)
                and p["s"] in (1, 2) and p["p"] in (None, p["k"] // 2) and cin % 32 == 0 and (p["s"] == 1 or (h % 2 == 0 and w % 2 == 0)))

    def linear(self, x: View, w, b, act="none", out=None, res=None):
        """x: token matrix as a (1,1,T,C) view; w: (out,in)."""
        if out is None:
            out = self.new(1, 1, x.W, w.shape[0])
        self.conv(x, w.float().view(w.shape[0], w.shape[1], 1, 1), b.float(), 1, 1, 1, act, out, res)
        return out

    # ---- modules -----------------------------------------------------------------------------------------
    def bottleneck(self, x, pfx, shortcut, g, out, k=(3, 3), e=1.0):
        """block.py:343-356."""
        mid = self.conv_bn(x, f"{pfx}.cv1", k[0])
        add = shortcut and x.C == out.C
        self.conv_bn(mid, f"{pfx}.cv2", k[1], 1, g, out=out, res=x if add else None)

    def c2f_fuses_cv2(self, p, c2, H, W, c3k2=False):
        """True when C2f.cv2 runs inside the launch of the block's last Bottleneck conv (ysod_conv_tc_set_b2b_cat)."""
        return (self.m.fuse_b2b and self.m.c2f_cat and not c3k2 and p["n"] == 1 and int(c2 * p["e"]) == 32 and c2 == 64 and p["g"] == 1
                and self.m.use_tc and self.code == _lib.BF16 and H >= 16 and W >= 8
                and H * W / (-(-H // 16) * -(-W // 8) * 128.0) >= 0.75)   # the halo (16 x 8 tile) plan will be chosen

    def c2f(self, x, P, p, c2, out, c3k2=False):
        """block.py:233-248 (C2f) / :733-741 (C3k2): cv1 -> chunk(2) -> n blocks -> cat -> cv2, with the cat buffer
        written slice by slice."""
        n = p["n"]
        c = int(c2 * p["e"])
        li = int(P.split(".")[1])
        fuse_cv2 = self.c2f_fuses_cv2(p, c2, out.H, out.W, c3k2)   # (x is None when cv1 ran inside its producer)
        if li in getattr(self, "_c2f_pre", {}):
            cat = self._c2f_pre[li]          # cv1 already ran inside the producer conv's launch
        else:
            # (fused cv2: the Bottleneck output never reaches HBM, so the buffer holds the cv1 output alone -- full 128 B lines per pixel)
            cat = self.new(out.N, out.H, out.W, (2 if fuse_cv2 else 2 + n) * c)
            self.conv_bn(x, f"{P}.cv1", out=cat.slice(0, 2 * c))
        if fuse_cv2:
            # cv2 (1x1, 96 -> 64) runs inside the launch of the Bottleneck's second 3x3 conv (ysod_conv_tc_set_b2b_cat): its input
            # cat(cv1 output, bottleneck output) = the 64-channel cv1 slice (TMA tile) + the staged 32-channel tile
            src = cat.slice(c, 2 * c)
            mid = self.conv_bn(src, f"{P}.m.0.cv1", 3)
            dummy = cat.slice(c, 2 * c)     # never written: the plan's output tensor map needs a valid address
            self.conv_bn(mid, f"{P}.m.0.cv2", 3, out=dummy, res=src if p["shortcut"] else None, no_duo=True)   # the cat fusion stages 64 B rows
            w2, b2 = self.folded(f"{P}.cv2")
            w2d, b2d = self.dev_t(w2.view(64, 3 * c), self.dt), self.dev_t(b2)
            self.call("ysod_conv_tc_set_b2b_cat", self.tc_handles[-1], cat.ptr(), cat.cs, _lib.ptr(w2d), _lib.ptr(b2d), _lib.ACT["silu"],
                      out.ptr(), out.cs)
            fl2 = 2.0 * cat.N * cat.H * cat.W * 64 * 3 * c
            self.op_flops[-1] += fl2
            self.tc_flops += fl2
            self.op_desc[-1] += " +C2f.cv2 96->64 over cat (b2b)"
            return
        for j in range(n):
            src, dst = cat.slice((1 + j) * c, (2 + j) * c), cat.slice((2 + j) * c, (3 + j) * c)
            if c3k2 and p["c3k"]:
                self.c3k(src, f"{P}.m.{j}", 2, p["shortcut"], p["g"], dst)
            elif c3k2:
                mid = self.conv_bn(src, f"{P}.m.{j}.cv1", 3)          # Bottleneck(c, c, shortcut, g) default e=0.5
                self.conv_bn(mid, f"{P}.m.{j}.cv2", 3, 1, p["g"], out=dst, res=src if p["shortcut"] else None)
            else:
                self.bottleneck(src, f"{P}.m.{j}", p["shortcut"], p["g"], dst)
        self.conv_bn(cat, f"{P}.cv2", out=out)

    def c3k(self, x, P, n, shortcut, g, out):
        """block.py:744-752 / :258-273: cv3(cat(m(cv1(x)), cv2(x)))."""
        c_ = self.m.sd[f"{P}.cv1.conv.weight"].shape[0]
        cat = self.new(x.N, x.H, x.W, 2 * c_)
        cur = self.conv_bn(x, f"{P}.cv1") if n > 0 else None
        for j in range(n):
            dst = cat.slice(0, c_) if j == n - 1 else self.new(x.N, x.H, x.W, c_)
            self.bottleneck(cur, f"{P}.m.{j}", shortcut, g, dst)
            cur = dst
        self.conv_bn(x, f"{P}.cv2", out=cat.slice(c_, 2 * c_))
        self.conv_bn(cat, f"{P}.cv3", out=out)

    def mha(self, tokens_norm: View, P, heads, batch, L, res: View = None):
        """nn.MultiheadAttention self-attention (batch_first): packed in_proj -> core -> out_proj (+ residual)."""
        sd = self.m.sd
        E = tokens_norm.C
        qkv = self.linear(tokens_norm, sd[f"{P}.in_proj_weight"], sd[f"{P}.in_proj_bias"])
        ao = self.new(1, 1, tokens_norm.W, E)
        D = E // heads
        esz = 1
        self.emit("ysod_mha_core_ex", qkv.ptr(), qkv.slice(E, 2 * E).ptr(), qkv.slice(2 * E, 3 * E).ptr(), self.code, batch, L, heads, D,
                  3 * E, 3 * E, 3 * E, L * 3 * E * esz, L * 3 * E * esz, L * 3 * E * esz, 1.0 / math.sqrt(D), ao.ptr(), E, L * E,
                  self.m.attn_impl, flops=4.0 * batch * heads * L * L * D, desc=f"mha b{batch} L{L} h{heads} d{D}")
        return self.linear(ao, sd[f"{P}.out_proj.weight"], sd[f"{P}.out_proj.bias"], res=res)

    def swin(self, x, P, p, out):
        """blocks_transformer.py:133-171."""
        sd = self.m.sd
        Cc, ws = x.C, p["window_size"]
        dww = sd[f"{P}.dw.weight"].float()
        A = f"{P}.window_attn"
        if (self.m.use_tc and self.code == _lib.BF16 and Cc == 64 and p["num_heads"] == 2 and ws == 7 and x.H > ws and x.W > ws
                and self.m.fuse_swin):
            # P2-level block: the whole SwinBlock is one kernel (csrc/swin_fused.cu)
            g, b = sd[f"{P}.bn.weight"].float(), sd[f"{P}.bn.bias"].float()
            sc = g / torch.sqrt(sd[f"{P}.bn.running_var"].float() + BN_EPS)
            pw = sd[f"{P}.pw.weight"].float().view(64, 64) * sc.view(-1, 1)
            # what is linear is folded on the host (swin_fused.cu): LayerNorm affine parts into the linear layers that follow them,
            # log2(e) / sqrt(head_dim) into the Q rows (the kernel's softmax exponentiates with ex2)
            g1, be1 = sd[f"{A}.norm1.weight"].float(), sd[f"{A}.norm1.bias"].float()
            g2, be2 = sd[f"{A}.norm2.weight"].float(), sd[f"{A}.norm2.bias"].float()
            Wi, bi = sd[f"{A}.attn.in_proj_weight"].float(), sd[f"{A}.attn.in_proj_bias"].float()
            W1, bm1 = sd[f"{A}.mlp.0.weight"].float(), sd[f"{A}.mlp.0.bias"].float()
            bi = bi + Wi @ be1
            Wi = Wi * g1.view(1, -1)
            qs = math.log2(math.e) / math.sqrt(32.0)
            Wi = torch.cat([Wi[:64] * qs, Wi[64:]])
            bi = torch.cat([bi[:64] * qs, bi[64:]])
            bm1 = bm1 + W1 @ be2
            W1 = W1 * g2.view(1, -1)
            # the key bias shifts every score of a query equally (softmax-invariant); the value bias passes through the attention average
            # (rows of softmax sum to 1) and lands in out_proj's bias: neither is added per token (ysod.h ysod_swin64_tc)
            Wo, bo = sd[f"{A}.attn.out_proj.weight"].float(), sd[f"{A}.attn.out_proj.bias"].float()
            bo = bo + Wo @ bi[128:]
            bi = torch.cat([bi[:64], torch.zeros(128)])
            wb = torch.cat([dww.view(64, 3, 3).permute(1, 2, 0).reshape(-1), Wi.reshape(-1),
                            sd[f"{A}.attn.out_proj.weight"].float().reshape(-1), W1.reshape(-1),
                            sd[f"{A}.mlp.2.weight"].float().reshape(-1), pw.reshape(-1)])
            pf = torch.cat([torch.ones(64), torch.zeros(64), bi,
                            bo, torch.ones(64), torch.zeros(64),
                            bm1, sd[f"{A}.mlp.2.bias"].float(),
                            b - sd[f"{P}.bn.running_mean"].float() * sc])
            assert wb.numel() == 37440 and pf.numel() == 768
            wbd, pfd = self.dev_t(wb, self.dt), self.dev_t(pf)
            T = x.N * (-(-x.H // ws)) * (-(-x.W // ws)) * ws * ws
            # swin_impl 0: tcgen05 / TMEM kernel (swin_tc.cu); 1: the mma.sync kernel (swin_fused.cu), the A/B baseline
            kern = "ysod_swin64_fused" if self.m.swin_impl == 1 else "ysod_swin64_tc"
            self.emit(kern, x.ptr(), x.N, x.H, x.W, x.cs, _lib.ptr(wbd), _lib.ptr(pfd), out.ptr(), out.cs, ws, 2,
                      flops=2.0 * T * (64 * 192 + 64 * 64 + 2 * 64 * 128 + 64 * 64 + 2 * 49 * 64),
                      desc=f"swin64 fused ({'mma.sync' if self.m.swin_impl == 1 else 'tcgen05'}) @{x.H}x{x.W}")
            return
        heads = p["num_heads"]
        D = Cc // heads
        if (self.m.swin_nhwc and self.m.use_tc and self.code == _lib.BF16 and Cc in (256, 512) and D in (32, 64) and ws * ws <= 64
                and x.H > ws and x.W > ws):
            # Tokens stay in NHWC pixel order (csrc/swin_nhwc.cu): dw 3x3 + LayerNorm 1 in one pass, the linears as 1x1 convs over the
            # N*H*W real pixels, the attention core reads / writes through the window -> pixel map (zero-padded window tokens enter as
            # the constant key / value of LayerNorm(0) = beta); no window_partition / window_reverse copies.
            g1, b1 = sd[f"{A}.norm1.weight"].float(), sd[f"{A}.norm1.bias"].float()
            y, yn = self.new(x.N, x.H, x.W, Cc), self.new(x.N, x.H, x.W, Cc)
            wdw = self.dev_t(dww.view(Cc, 3, 3).permute(1, 2, 0).contiguous())   # [3][3][C] fp32
            self.emit("ysod_dwconv3_ln", x.ptr(), x.N, x.H, x.W, Cc, x.cs, _lib.ptr(wdw), _lib.ptr(self.dev_t(g1)), _lib.ptr(self.dev_t(b1)),
                      LN_EPS, y.ptr(), y.cs, yn.ptr(), yn.cs, desc=f"dw3x3 + LayerNorm {Cc} @{x.H}x{x.W}")
            T = x.N * x.H * x.W
            tok = lambda v: View(v.buf, v.off, 1, 1, T, v.C, v.cs)
            Wi, bi = sd[f"{A}.attn.in_proj_weight"].float(), sd[f"{A}.attn.in_proj_bias"].float()
            qkv = self.linear(tok(yn), Wi, bi)
            # key / value of a zero-padded token: in_proj(LayerNorm(0)) with the operands rounded as the kernels round them
            lp = self.dt
            beta_lp = b1.to(lp).float()
            kv_pad = (Wi[Cc:].to(lp).float() @ beta_lp + bi[Cc:]).to(lp)
            kvd = self.dev_t(kv_pad, lp)
            ao = self.new(1, 1, T, Cc)
            self.emit("ysod_mha_window_nhwc", qkv.ptr(), qkv.slice(Cc, 2 * Cc).ptr(), qkv.slice(2 * Cc, 3 * Cc).ptr(), 3 * Cc, x.N, x.H, x.W, ws,
                      heads, D, _lib.ptr(kvd), _lib.ptr(kvd, Cc), 1.0 / math.sqrt(D), ao.ptr(), Cc,
                      flops=4.0 * x.N * (-(-x.H // ws)) * (-(-x.W // ws)) * heads * (ws * ws) ** 2 * D, desc=f"window attention (NHWC gather) h{heads} d{D}")
            w2 = self.linear(ao, sd[f"{A}.attn.out_proj.weight"], sd[f"{A}.attn.out_proj.bias"], res=tok(y))
            n2 = self.new(1, 1, T, Cc)
            g2, b2 = self.dev_t(sd[f"{A}.norm2.weight"]), self.dev_t(sd[f"{A}.norm2.bias"])
            self.emit("ysod_layernorm", w2.ptr(), self.code, T, Cc, w2.cs, _lib.ptr(g2), _lib.ptr(b2), LN_EPS, n2.ptr(), n2.cs)
            hmid = self.linear(n2, sd[f"{A}.mlp.0.weight"], sd[f"{A}.mlp.0.bias"], act="gelu")
            w3 = self.linear(hmid, sd[f"{A}.mlp.2.weight"], sd[f"{A}.mlp.2.bias"], res=w2)
            r = View(w3.buf, w3.off, x.N, x.H, x.W, Cc, w3.cs)
            pw = sd[f"{P}.pw.weight"].float()
            g, b = sd[f"{P}.bn.weight"].float(), sd[f"{P}.bn.bias"].float()
            sc = g / torch.sqrt(sd[f"{P}.bn.running_var"].float() + BN_EPS)
            self.conv(r, pw * sc.view(-1, 1, 1, 1), b - sd[f"{P}.bn.running_mean"].float() * sc, 1, 1, 1, "silu", out, res=x)
            return
        y = self.new(x.N, x.H, x.W, Cc)
        self.conv(x, dww, torch.zeros(Cc), 3, 1, Cc, "none", y)
        wh, ww = min(ws, x.H), min(ws, x.W)
        if x.H <= ws and x.W <= ws:
            nWh = nWw = 1
        else:
            nWh, nWw = -(-x.H // wh), -(-x.W // ww)
        T = x.N * nWh * nWw * wh * ww
        raw, nrm = self.new(1, 1, T, Cc), self.new(1, 1, T, Cc)
        g1, b1 = self.dev_t(sd[f"{A}.norm1.weight"]), self.dev_t(sd[f"{A}.norm1.bias"])
        self.emit("ysod_window_partition_ln", y.ptr(), self.code, x.N, x.H, x.W, Cc, y.cs, wh, ww, nWh, nWw, _lib.ptr(g1), _lib.ptr(b1),
                  LN_EPS, raw.ptr(), nrm.ptr(), Cc)
        w2 = self.mha(nrm, f"{A}.attn", p["num_heads"], x.N * nWh * nWw, wh * ww, res=raw)
        n2 = self.new(1, 1, T, Cc)
        g2, b2 = self.dev_t(sd[f"{A}.norm2.weight"]), self.dev_t(sd[f"{A}.norm2.bias"])
        self.emit("ysod_layernorm", w2.ptr(), self.code, T, Cc, w2.cs, _lib.ptr(g2), _lib.ptr(b2), LN_EPS, n2.ptr(), n2.cs)
        h = self.linear(n2, sd[f"{A}.mlp.0.weight"], sd[f"{A}.mlp.0.bias"], act="gelu")
        w3 = self.linear(h, sd[f"{A}.mlp.2.weight"], sd[f"{A}.mlp.2.bias"], res=w2)
        r = self.new(x.N, x.H, x.W, Cc)
        self.emit("ysod_window_reverse", w3.ptr(), self.code, w3.cs, x.N, x.H, x.W, Cc, wh, ww, nWh, nWw, r.ptr(), r.cs)
        # pw 1x1 (no bias) -> BN -> SiLU -> + identity
        pw = sd[f"{P}.pw.weight"].float()
        g, b = sd[f"{P}.bn.weight"].float(), sd[f"{P}.bn.bias"].float()
        s = g / torch.sqrt(sd[f"{P}.bn.running_var"].float() + BN_EPS)
        self.conv(r, pw * s.view(-1, 1, 1, 1), b - sd[f"{P}.bn.running_mean"].float() * s, 1, 1, 1, "silu", out, res=x)

    def fold_bn(self, w, bn):
        """conv weight (Cout, ...) with the BatchNorm2d at state_dict prefix `bn` folded in: (w', bias)."""
        sd = self.m.sd
        s = sd[f"{bn}.weight"].float() / torch.sqrt(sd[f"{bn}.running_var"].float() + BN_EPS)
        return w.float() * s.view(-1, *([1] * (w.dim() - 1))), sd[f"{bn}.bias"].float() - sd[f"{bn}.running_mean"].float() * s

    def mamba(self, x, P, p, out):
        """blocks_mamba.py:105-111,167-236 with the GLU fallback (:84-103): in_proj -> avg_pool(r) -> pw1 -> sigmoid(g)*a -> dw3x3+BN+SiLU
        -> pw2 -> [nearest upsample -> out_proj -> + x]. out_proj (1x1 + BN + SiLU) is pointwise, so it runs on the pooled map and the
        upsample is fused with the residual add: identical values, a quarter of the work."""
        sd = self.m.sd
        ch, r = p["c_hidden"], p["reduction"]
        hid = sd[f"{P}.fallback.dw.weight"].shape[0]
        w, b = self.fold_bn(sd[f"{P}.in_proj.0.weight"], f"{P}.in_proj.1")
        y = self.new(x.N, x.H, x.W, ch)
        self.conv(x, w, b, 1, 1, 1, "silu", y)
        if r > 1:
            yp = self.new(x.N, x.H // r, x.W // r, ch)
            self.emit("ysod_avgpool2d", y.ptr(), self.code, y.N, y.H, y.W, ch, y.cs, r, yp.ptr(), yp.cs, desc=f"avgpool{r}")
            y = yp
        t = self.new(y.N, y.H, y.W, 2 * hid)
        self.conv(y, sd[f"{P}.fallback.pw1.weight"].float(), torch.zeros(2 * hid), 1, 1, 1, "none", t)
        gl = self.new(y.N, y.H, y.W, hid)
        self.emit("ysod_glu", t.ptr(), self.code, y.N * y.H * y.W, hid, t.cs, gl.ptr(), gl.cs, desc="GLU gate")
        w, b = self.fold_bn(sd[f"{P}.fallback.dw.weight"], f"{P}.fallback.bn")
        d = self.new(y.N, y.H, y.W, hid)
        self.conv(gl, w, b, 3, 1, hid, "silu", d)
        q = self.new(y.N, y.H, y.W, ch)
        self.conv(d, sd[f"{P}.fallback.pw2.weight"].float(), torch.zeros(ch), 1, 1, 1, "none", q)
        w, b = self.fold_bn(sd[f"{P}.out_proj.0.weight"], f"{P}.out_proj.1")
        o2 = self.new(y.N, y.H, y.W, x.C)
        self.conv(q, w, b, 1, 1, 1, "silu", o2)
        self.emit("ysod_upsample_add", o2.ptr(), self.code, x.N, o2.H, o2.W, x.C, o2.cs, x.ptr(), x.cs, x.H, x.W, out.ptr(), out.cs,
                  desc="nearest upsample + residual")

    def a2attn(self, x, P, p, out):
        """a2_attn.py:35-69."""
        sd = self.m.sd
        Cc, na = x.C, p["num_areas"]
        xp = self.conv_bn(x, f"{P}.proj")
        pooled = self.new(x.N, na, x.W, Cc)
        self.emit("ysod_adaptive_pool_rows", xp.ptr(), self.code, x.N, x.H, x.W, Cc, xp.cs, na, pooled.ptr(), pooled.cs)
        T = x.N * na * x.W
        seq = View(pooled.buf, pooled.off, 1, 1, T, Cc, Cc)
        sn = self.new(1, 1, T, Cc)
        g, b = self.dev_t(sd[f"{P}.layer_norm.weight"]), self.dev_t(sd[f"{P}.layer_norm.bias"])
        self.emit("ysod_layernorm", seq.ptr(), self.code, T, Cc, Cc, _lib.ptr(g), _lib.ptr(b), LN_EPS, sn.ptr(), Cc)
        ao = self.mha(sn, f"{P}.attention", p["num_heads"], x.N, na * x.W)
        up = self.new(x.N, x.H, x.W, Cc)
        self.emit("ysod_bilinear_rows", ao.ptr(), self.code, x.N, na, x.W, Cc, ao.cs, x.H, up.ptr(), up.cs)
        self.conv_bn(up, f"{P}.out_proj", out=out, res=x if out.C == x.C else None)   # a2_attn.py:64-69: identity only if c2 == c1

    def ablock(self, x, P, heads, area, out):
        """block.py:1367-1416 ABlock with AAttn (:1298-1365, manual-softmax semantics)."""
        Cc = x.C
        A = f"{P}.attn"
        qk = self.conv_bn(x, f"{A}.qk", act=False)
        v = self.conv_bn(x, f"{A}.v", act=False)
        L = x.H * x.W
        if area > 1:
            assert L % area == 0, "AAttn: H*W must be divisible by area"
            L //= area
        batch = x.N * max(area, 1)
        D = Cc // heads
        ao = self.new(x.N, x.H, x.W, Cc)
        self.emit("ysod_mha_core_ex", qk.ptr(), qk.slice(Cc, 2 * Cc).ptr(), v.ptr(), self.code, batch, L, heads, D, 2 * Cc, 2 * Cc, Cc,
                  L * 2 * Cc, L * 2 * Cc, L * Cc, D ** -0.5, ao.ptr(), Cc, L * Cc, self.m.attn_impl,
                  flops=4.0 * batch * heads * L * L * D, desc=f"area attention b{batch} L{L} h{heads} d{D}")
        s = self.conv_bn(v, f"{A}.pe", 5, 1, Cc, act=False, res=ao, pad=2)     # pe(v) + attention output
        x1 = self.conv_bn(s, f"{A}.proj", act=False, res=x)                    # x + proj(...)
        h = self.conv_bn(x1, f"{P}.mlp.0")
        self.conv_bn(h, f"{P}.mlp.1", act=False, out=out, res=x1)

    def a2c2f(self, x, P, p, c2, out):
        """block.py:1418-1472."""
        if p["a2"] and p["residual"]:
            raise NotImplementedError("A2C2f residual/gamma variant (l/x scales) is outside the supported configs")
        n = p["n"]
        c_ = int(c2 * p["e"])
        cat = self.new(x.N, x.H, x.W, (1 + n) * c_)
        self.conv_bn(x, f"{P}.cv1", out=cat.slice(0, c_))
        for j in range(n):
            src, dst = cat.slice(j * c_, (j + 1) * c_), cat.slice((j + 1) * c_, (j + 2) * c_)
            if p["a2"]:
                mid = self.new(x.N, x.H, x.W, c_)
                self.ablock(src, f"{P}.m.{j}.0", c_ // 32, p["area"], mid)
                self.ablock(mid, f"{P}.m.{j}.1", c_ // 32, p["area"], dst)
            else:
                self.c3k(src, f"{P}.m.{j}", 2, p["shortcut"], p["g"], dst)
        self.conv_bn(cat, f"{P}.cv2", out=out)

    def sppf(self, x, P, p, out):
        c_ = x.C // 2
        cat = self.new(x.N, x.H, x.W, 4 * c_)
        self.conv_bn(x, f"{P}.cv1", out=cat.slice(0, c_))
        self.emit("ysod_sppf_pool", cat.ptr(), self.code, x.N, x.H, x.W, c_, cat.cs, p["k"], cat.slice(c_, 2 * c_).ptr(),
                  cat.slice(2 * c_, 3 * c_).ptr(), cat.slice(3 * c_, 4 * c_).ptr(), cat.cs)
        self.conv_bn(cat, f"{P}.cv2", out=out)

    def _splits(self, N, HW):
        return max(1, min(HW // 256 if HW >= 256 else 1, -(-1184 // N)))   # ~8 CTAs (2048 threads) per SM, >= 256 pixels per split

    def se(self, x, P, out):
        sd = self.m.sd
        Cc, HW = x.C, x.H * x.W
        hid = sd[f"{P}.fc1.weight"].shape[0]
        S = self._splits(x.N, HW)
        pre = getattr(self, "_stem_psum", None)
        pre = pre if (pre is not None and pre[0].buf is x.buf and pre[0].off == x.off and pre[0].C == x.C) else None
        psum, gate = (pre[1] if pre else self.f32(x.N, S, Cc)), self.f32(x.N, Cc)
        w1, b1 = self.dev_t(sd[f"{P}.fc1.weight"].reshape(hid, Cc)), self.dev_t(sd[f"{P}.fc1.bias"])
        w2, b2 = self.dev_t(sd[f"{P}.fc2.weight"].reshape(Cc, hid)), self.dev_t(sd[f"{P}.fc2.bias"])
        if pre:     # the producing stem kernel already wrote the pooling partials (ysod_stem_mma_gap)
            self.emit("ysod_se_gate", _lib.ptr(psum), x.N, pre[2], HW, Cc, _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), hid, _lib.ptr(gate))
        elif self.m.fuse_gate:
            cnt = torch.zeros(x.N, device=self.dev, dtype=torch.int32)
            self.keep.append(cnt)
            self.emit("ysod_gap_gate", x.ptr(), self.code, x.N, HW, Cc, x.cs, S, _lib.ptr(psum), None, _lib.ptr(cnt), 0, _lib.ptr(w1),
                      _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), hid, _lib.ptr(gate), desc="GAP + SE gate")
        else:
            self.emit("ysod_gap_partial", x.ptr(), self.code, x.N, HW, Cc, x.cs, S, _lib.ptr(psum), None)
            self.emit("ysod_se_gate", _lib.ptr(psum), x.N, S, HW, Cc, _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2), hid, _lib.ptr(gate))
        if out is None:
            return gate          # `x * gate` is folded into the consumer conv's weights
        self.emit("ysod_scale_channels", x.ptr(), self.code, x.N, HW, Cc, x.cs, _lib.ptr(gate), out.ptr(), out.cs)

    def cbam(self, x, P, out):
        sd = self.m.sd
        Cc, HW = x.C, x.H * x.W
        hid = sd[f"{P}.channel_attention.fc.0.weight"].shape[0]
        S = self._splits(x.N, HW)
        psum, pmax, gate = self.f32(x.N, S, Cc), self.f32(x.N, S, Cc), self.f32(x.N, Cc)
        stats = self.f32(x.N, HW, 2)
        w1 = self.dev_t(sd[f"{P}.channel_attention.fc.0.weight"].reshape(hid, Cc))
        w2 = self.dev_t(sd[f"{P}.channel_attention.fc.2.weight"].reshape(Cc, hid))
        wsp = self.dev_t(sd[f"{P}.spatial_attention.conv1.weight"].reshape(2, 7, 7))
        if self.m.fuse_gate:
            cnt = torch.zeros(x.N, device=self.dev, dtype=torch.int32)
            self.keep.append(cnt)
            self.emit("ysod_gap_gate", x.ptr(), self.code, x.N, HW, Cc, x.cs, S, _lib.ptr(psum), _lib.ptr(pmax), _lib.ptr(cnt), 1,
                      _lib.ptr(w1), None, _lib.ptr(w2), None, hid, _lib.ptr(gate), desc="GAP + CBAM channel gate")
        else:
            self.emit("ysod_gap_partial", x.ptr(), self.code, x.N, HW, Cc, x.cs, S, _lib.ptr(psum), _lib.ptr(pmax))
            self.emit("ysod_cbam_gate", _lib.ptr(psum), _lib.ptr(pmax), x.N, S, HW, Cc, _lib.ptr(w1), _lib.ptr(w2), hid, _lib.ptr(gate))
        c8 = Cc // 8
        if self.m.fuse_cbam and c8 <= 32 and (c8 & (c8 - 1)) == 0:
            # statistics + 7x7 conv + apply in one pass over the map (2 reads + 1 write for the whole block)
            self.emit("ysod_cbam_spatial", x.ptr(), self.code, x.N, x.H, x.W, Cc, x.cs, _lib.ptr(gate), _lib.ptr(wsp), 7, out.ptr(), out.cs,
                      desc="CBAM spatial (stats + 7x7 + apply)")
            return
        self.emit("ysod_cbam_stats", x.ptr(), self.code, x.N, HW, Cc, x.cs, _lib.ptr(gate), _lib.ptr(stats))
        self.emit("ysod_cbam_apply", x.ptr(), self.code, x.N, x.H, x.W, Cc, x.cs, _lib.ptr(gate), _lib.ptr(stats), _lib.ptr(wsp), 7,
                  out.ptr(), out.cs)

    def ca(self, x, P, out):
        sd = self.m.sd
        Cc = x.C
        mip = sd[f"{P}.conv1.weight"].shape[0]
        s = sd[f"{P}.bn1.weight"].float() / torch.sqrt(sd[f"{P}.bn1.running_var"].float() + BN_EPS)
        w1 = self.dev_t(sd[f"{P}.conv1.weight"].float().reshape(mip, Cc) * s.view(-1, 1))
        b1 = self.dev_t((sd[f"{P}.conv1.bias"].float() - sd[f"{P}.bn1.running_mean"].float()) * s + sd[f"{P}.bn1.bias"].float())
        wh, bh = self.dev_t(sd[f"{P}.conv_h.weight"].reshape(Cc, mip)), self.dev_t(sd[f"{P}.conv_h.bias"])
        ww, bw = self.dev_t(sd[f"{P}.conv_w.weight"].reshape(Cc, mip)), self.dev_t(sd[f"{P}.conv_w.bias"])
        pooled, att = self.f32(x.N, x.H + x.W, Cc), self.f32(x.N, x.H + x.W, Cc)
        nws = int(self.lib.ysod_ca_pool_workspace_floats(x.N, x.H, x.W, Cc)) if self.m.ca_single_pass else 0
        ws = self.f32(nws) if nws > 0 else None
        self.emit("ysod_ca_pool", x.ptr(), self.code, x.N, x.H, x.W, Cc, x.cs, _lib.ptr(pooled), _lib.ptr(ws) if ws is not None else None)
        self.emit("ysod_ca_gate", _lib.ptr(pooled), x.N, x.H, x.W, Cc, mip, _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(wh), _lib.ptr(bh),
                  _lib.ptr(ww), _lib.ptr(bw), _lib.ptr(att))
        self.emit("ysod_ca_apply", x.ptr(), self.code, x.N, x.H, x.W, Cc, x.cs, _lib.ptr(att), out.ptr(), out.cs)

    def detect(self, xs, P, p, src_layers):
        """head.py:64-131: per-level cv2/cv3 stacks -> raw maps (fp32 NHWC, stride RAW_CS) -> fused decode."""
        sd = self.m.sd
        nc = p["nc"]
        no = 64 + nc
        raw_cs = max(RAW_CS, (no + 7) // 8 * 8)
        A = sum(v.H * v.W for v in xs)
        y = torch.empty((self.B, 4 + nc, A), device=self.dev, dtype=torch.float32)
        self.keep.append(y)
        raws = []
        a_off = 0
        nl = len(xs)
        for i, x in enumerate(xs):
            # Each level's head chain only depends on its own input map: all but the last level run on their own stream lane,
            # forked right after the layer that produced the input (the P2 head overlaps the rest of the neck), joined at the end.
            branch = self.m.multi_stream and i < nl - 1
            if branch:
                self._lane = 1 + i
                self.n_lanes = max(self.n_lanes, 2 + i)
                self.sched.append(("wait", f"L{src_layers[i]}", self._lane))
            fused_tail = (p["legacy"] and self.m.use_tc and self.code == _lib.BF16 and self.m.fuse_decode and no <= 128
                          and (sd[f"{P}.cv2.{i}.0.conv.weight"].shape[0] + sd[f"{P}.cv3.{i}.1.conv.weight"].shape[0]) % 32 == 0)
            skip_raw = not self.want_raw and fused_tail   # predict path: `y` is written by the tail conv's epilogue, nothing reads raw
            raw = None if skip_raw else self.new(x.N, x.H, x.W, raw_cs, dtype=torch.float32, zero=True)
            decoded = False
            if p["legacy"] and self.m.use_tc and self.code == _lib.BF16:
                # cv2[i][0] and cv3[i][0] are 3x3 convs on the same input: run them as ONE conv with N = c2 + c3 output channels
                # (the activation tile is fetched once, and N = 128 MMAs run at 93 % of the tensor rate vs 60 % for N = 64)
                w2, b2 = self.folded(f"{P}.cv2.{i}.0")
                w3, b3 = self.folded(f"{P}.cv3.{i}.0")
                c2n = w2.shape[0]
                both = self.new(x.N, x.H, x.W, c2n + w3.shape[0])
                self.conv(x, torch.cat([w2, w3], 0), torch.cat([b2, b3], 0), 3, 1, 1, "silu", both)
                c3n = sd[f"{P}.cv3.{i}.1.conv.weight"].shape[0]
                if (c2n + c3n) % 32 == 0 and no <= 128:
                    # cv2[i][2] and cv3[i][2] (1x1, bias, no act) as ONE block-diagonal 1x1 conv [64 + nc][c2 + c3] over the two
                    # branch outputs stored side by side: the pair is HBM-bound, so the zero half of the weights is free and the
                    # raw map is written by one launch (full 74-channel rows instead of two partial slices)
                    if (self.m.fuse_b2b and self.m.fuse_decode and c2n == 64 and c3n == 64 and nc <= 64
                            and sd[f"{P}.cv2.{i}.1.conv.weight"].shape[0] == 64):
                        # back-to-back GEMM: each branch's 3x3 conv feeds its 1x1 head conv from the staged tile and decodes its share of
                        # y in the second epilogue (ysod_conv_tc_set_b2b): no `ac` buffer, no tail launch, raw map only on request
                        for br, (lo, hi, kind, n2, coff) in enumerate([(0, c2n, 1, 64, 0), (c2n, both.C, 2, (nc + 15) // 16 * 16, 64)]):
                            name = f"{P}.cv{2 + br}.{i}"
                            dummy = View(both.buf, both.off, x.N, x.H, x.W, 64, both.cs)   # never written (the tensor map needs an address)
                            self.conv_bn(both.slice(lo, hi), f"{name}.1", 3, out=dummy)
                            w2 = torch.zeros((n2, 64), dtype=torch.float32)
                            b2 = torch.zeros(n2, dtype=torch.float32)
                            wsrc = sd[f"{name}.2.weight"].float().view(-1, 64)
                            w2[:wsrc.shape[0]] = wsrc
                            b2[:wsrc.shape[0]] = sd[f"{name}.2.bias"].float()
                            w2d, b2d = self.dev_t(w2, self.dt), self.dev_t(b2)
                            self.call("ysod_conv_tc_set_b2b", self.tc_handles[-1], _lib.ptr(w2d), _lib.ptr(b2d), n2, kind, _lib.ptr(y), A, a_off, nc,
                                      float(self.m.stride_list[i]), raw.ptr() if raw is not None else None, raw_cs, coff)
                            fl2 = 2.0 * x.N * x.H * x.W * 64 * (64 if kind == 1 else nc)
                            self.op_flops[-1] += fl2
                            self.tc_flops += fl2
                            self.op_desc[-1] += " +1x1 head conv +decode (b2b)"
                        decoded = True
                        a = c = None
                        a_off += x.H * x.W
                        if raw is not None:
                            raws.append(raw.slice(0, no))
                        if branch:
                            self.sched.append(("record", f"det{i}", self._lane))
                            self._lane = 0
                        continue
                    ac = self.new(x.N, x.H, x.W, c2n + c3n)
                    self.conv_bn(both.slice(0, c2n), f"{P}.cv2.{i}.1", 3, out=ac.slice(0, c2n))
                    self.conv_bn(both.slice(c2n, both.C), f"{P}.cv3.{i}.1", 3, out=ac.slice(c2n, c2n + c3n))
                    wt = torch.zeros((no, c2n + c3n, 1, 1), dtype=torch.float32)
                    wt[:64, :c2n] = sd[f"{P}.cv2.{i}.2.weight"].float()
                    wt[64:, c2n:] = sd[f"{P}.cv3.{i}.2.weight"].float()
                    bt = torch.cat([sd[f"{P}.cv2.{i}.2.bias"].float(), sd[f"{P}.cv3.{i}.2.bias"].float()])
                    if skip_raw:   # the tensor map still needs a valid address; YSOD_CONV_NO_STORE never writes through it
                        dummy = View(y, 0, x.N, x.H, x.W, no, raw_cs)
                        self.conv(ac, wt, bt, 1, 1, 1, "none", dummy, out_f32=True, no_store=True,
                                  alg_flops=2.0 * x.N * x.H * x.W * (64 * c2n + nc * c3n))
                    else:
                        self.conv(ac, wt, bt, 1, 1, 1, "none", raw.slice(0, no), out_f32=True,
                                  alg_flops=2.0 * x.N * x.H * x.W * (64 * c2n + nc * c3n))
                    a = c = None
                    if self.m.fuse_decode:
                        # ... and the level's DFL / dist2bbox / sigmoid decode runs in that conv's epilogue (no re-read of the raw map)
                        self.call("ysod_conv_tc_set_decode", self.tc_handles[-1], _lib.ptr(y), A, a_off, nc, float(self.m.stride_list[i]))
                        self.op_desc[-1] += " +decode"
                        decoded = True
                else:
                    a = self.conv_bn(both.slice(0, c2n), f"{P}.cv2.{i}.1", 3)
                    c = self.conv_bn(both.slice(c2n, both.C), f"{P}.cv3.{i}.1", 3)
            else:
                a = self.conv_bn(x, f"{P}.cv2.{i}.0", 3)
                a = self.conv_bn(a, f"{P}.cv2.{i}.1", 3)
                if p["legacy"]:
                    c = self.conv_bn(x, f"{P}.cv3.{i}.0", 3)
                    c = self.conv_bn(c, f"{P}.cv3.{i}.1", 3)
                else:
                    c = self.conv_bn(x, f"{P}.cv3.{i}.0.0", 3, 1, x.C)
                    c = self.conv_bn(c, f"{P}.cv3.{i}.0.1", 1)
                    c = self.conv_bn(c, f"{P}.cv3.{i}.1.0", 3, 1, c.C)
                    c = self.conv_bn(c, f"{P}.cv3.{i}.1.1", 1)
            if a is not None:
                self.conv(a, sd[f"{P}.cv2.{i}.2.weight"].float(), sd[f"{P}.cv2.{i}.2.bias"].float(), 1, 1, 1, "none", raw.slice(0, 64),
                          out_f32=True)
                self.conv(c, sd[f"{P}.cv3.{i}.2.weight"].float(), sd[f"{P}.cv3.{i}.2.bias"].float(), 1, 1, 1, "none",
                          raw.slice(64, 64 + nc), out_f32=True)
            if not decoded:
                self.emit("ysod_dfl_decode", raw.ptr(), _lib.F32, x.N, x.H, x.W, raw.cs, nc, 16, float(self.m.stride_list[i]), _lib.ptr(y), A,
                          a_off)
            a_off += x.H * x.W
            if raw is not None:
                raws.append(raw.slice(0, no))
            if branch:
                self.sched.append(("record", f"det{i}", self._lane))
                self._lane = 0
        for i in range(nl - 1):
            if self.m.multi_stream:
                self.sched.append(("wait", f"det{i}", 0))
        self.y = y
        self.raws = raws

    # ---- graph walk --------------------------------------------------------------------------------------
    def _build(self):
        spec = self.m.spec
        layers = spec.layers
        # shapes of every layer output
        shp = {}
        for L in layers:
            if L.i == 0:
                cin_hw = (self.H, self.W)
            src = L.f if isinstance(L.f, int) else L.f[0]
            src = L.i - 1 if src == -1 else src
            h, w = (self.H, self.W) if L.i == 0 else shp[src][:2]
            if L.type == "Conv":
                k, s = L.p["k"], L.p["s"]
                pd = k // 2 if L.p["p"] is None else L.p["p"]
                h, w = (h + 2 * pd - k) // s + 1, (w + 2 * pd - k) // s + 1
            elif L.type == "nn.Upsample":
                h, w = h * L.p["scale"], w * L.p["scale"]
            shp[L.i] = (h, w, L.c2)
        # homes: outputs that are produced directly inside a Concat buffer
        home: Dict[int, View] = {}
        cat_buf: Dict[int, View] = {}
        for L in layers:
            if L.type == "Concat":
                h, w, c = shp[L.i]
                cb = self.new(self.B, h, w, c)
                cat_buf[L.i] = cb
                c0 = 0
                for j, src in enumerate(L.f):
                    src = L.i - 1 if src == -1 else src
                    cj = L.c1[j]
                    if src not in home and layers[src].type not in ("Concat", "Detect"):
                        home[src] = cb.slice(c0, c0 + cj)
                    c0 += cj
        out = self.layer_out
        det_inputs = set(layers[-1].f) if layers[-1].type == "Detect" else set()
        # Conv -> nn.Upsample(2) pairs where the conv feeds nothing else (the neck's top-down path): the conv stores straight into
        # the upsampled map (TMA store with element stride 2, four phases) and the upsample launch disappears
        consumers: Dict[int, list] = {}
        for L in layers:
            for src in (L.f if isinstance(L.f, list) else [L.f]):
                consumers.setdefault(L.i - 1 if src == -1 else src, []).append(L.i)
        fold_up: Dict[int, int] = {}
        for L in layers:
            if L.type == "nn.Upsample" and L.p["scale"] == 2 and L.p.get("mode", "nearest") == "nearest" and L.i > 0:
                src = L.i - 1 if L.f == -1 else L.f
                S = layers[src]
                if (isinstance(src, int) and S.type == "Conv" and S.i > 0 and consumers.get(src) == [L.i] and src not in home
                        and src not in det_inputs):
                    ssrc = S.i - 1 if S.f == -1 else S.f
                    if self.m.fuse_upsample and self.tc_eligible(S, shp[ssrc][2], shp[ssrc][0], shp[ssrc][1]):
                        fold_up[src] = L.i
        # SE -> Conv pairs where the SE output feeds nothing else: the channel gate goes into per-image conv weights
        fold_se = set()
        for L in layers:
            if L.type in ("SE_Block", "SE") and L.i > 0 and self.m.fuse_se and L.i not in home and L.i not in det_inputs:
                cons = consumers.get(L.i, [])
                if len(cons) == 1 and layers[cons[0]].type == "Conv":   # (the consumer may itself store through a fused nn.Upsample)
                    Cn = layers[cons[0]]
                    src_hw = shp[L.i]
                    if self.tc_eligible(Cn, src_hw[2], src_hw[0], src_hw[1]):
                        fold_se.add(L.i)
        # Conv -> C2f pairs where the conv (64 output channels, tensor-core path) feeds nothing else and C2f.cv1 is 64 -> 64: cv1 runs as
        # the second MMA group of the conv's launch (ysod_conv_tc_set_b2b_conv); the conv's own output never reaches HBM
        fold_cv1: Dict[int, int] = {}
        self._c2f_pre: Dict[int, View] = {}
        for L in layers:
            if (L.type == "C2f" and self.m.fuse_b2b and isinstance(L.f, int) and self.m.use_tc and self.code == _lib.BF16):
                src = L.i - 1 if L.f == -1 else L.f
                S = layers[src]
                ssrc = S.i - 1 if S.f == -1 else S.f
                if (S.type == "Conv" and S.i > 0 and isinstance(S.f, int) and consumers.get(src) == [L.i] and src not in home and src not in det_inputs
                        and src not in fold_up and shp[src][2] == 64 and 2 * int(L.c2 * L.p["e"]) == 64
                        and self.tc_eligible(S, shp[ssrc][2], shp[ssrc][0], shp[ssrc][1])):
                    fold_cv1[src] = L.i
        for L in layers:
            P, p, t = f"model.{L.i}", L.p, L.type
            self._ctx = f"L{L.i}:{t}"
            if t == "Detect":
                self.detect([out[j] for j in L.f], P, p, list(L.f))
                continue
            if t == "Concat":
                cb = cat_buf[L.i]
                c0 = 0
                for j, src in enumerate(L.f):
                    src = L.i - 1 if src == -1 else src
                    cj = L.c1[j]
                    sv, dv = out[src], cb.slice(c0, c0 + cj)
                    if not (sv.buf is dv.buf and sv.off == dv.off):
                        self.emit("ysod_upsample_copy", sv.ptr(), self.code, sv.N, sv.H, sv.W, sv.C, sv.cs, 1, dv.ptr(), dv.cs)
                    c0 += cj
                out[L.i] = cb
                if L.i in det_inputs:
                    self.sched.append(("record", f"L{L.i}", 0))
                continue
            h, w, c = shp[L.i]
            src = L.i - 1 if L.f == -1 else L.f
            if t == "nn.Upsample" and fold_up.get(src) == L.i:
                continue                      # already written by its producer conv
            if L.i in fold_up:
                U = fold_up[L.i]
                uh, uw, uc = shp[U]
                uo = home[U] if U in home else self.new(self.B, uh, uw, uc)
                xin, gate = out[src], None
                if isinstance(xin, GatedView):
                    xin, gate = xin.base, xin.gate
                self.conv_bn(xin, P, p["k"], p["s"], p["g"], p["act"], out=uo, pad=p["p"], up2=True, gate=gate)
                out[U] = uo
                out[L.i] = View(uo.buf, uo.off, uo.N, h, w, c, uo.cs, sub=2)
                if U in det_inputs:
                    self.sched.append(("record", f"L{U}", 0))
                continue
            if L.i in fold_se:
                out[L.i] = GatedView(out[src], self.se(out[src], P, None))
                continue
            o = None if L.i in fold_cv1 else (home[L.i] if L.i in home else self.new(self.B, h, w, c))
            x = None if L.i == 0 else out[src]
            gate = None
            if isinstance(x, GatedView):
                x, gate = x.base, x.gate
            if t == "Conv":
                if L.i == 0:
                    wf, bf = self.folded(P)
                    k, s = p["k"], p["s"]
                    assert p["g"] == 1 and wf.shape[1] == 3, "stem must be a dense conv on a 3-channel image"
                    pd = k // 2 if p["p"] is None else p["p"]
                    co = wf.shape[0]
                    if (self.code == _lib.BF16 and self.m.use_tc and k == 3 and s == 2 and pd == 1 and co in (16, 32, 64)
                            and self.H % 2 == 0 and self.W % 2 == 0):
                        # tensor-core stem: weights as [Cout][32] bf16, column (r*3+s)*3+c, zero padded
                        wk = torch.zeros((co, 32), dtype=torch.float32)
                        wk[:, :27] = wf.permute(0, 2, 3, 1).reshape(co, 27)
                        wd, bd = self.dev_t(wk, self.dt), self.dev_t(bf)
                        self.img_indirect = True
                        nxt = layers[1] if len(layers) > 1 else None
                        if self.m.fuse_stem_gap and nxt is not None and nxt.type == "SE_Block" and nxt.f == -1:
                            # the SE block after the stem pools the stem's output: the stem kernel leaves the per-tile channel sums (SE_Block avg_pool)
                            S0 = -(-(self.W // 2) // 64) * -(-(self.H // 2) // 4)
                            ps0 = self.f32(self.B, S0, co)
                            self._stem_psum = (o, ps0, S0)
                            self.emit("ysod_stem_mma_gap", _lib.ptr(self.img_slot), (1 if self.src_u8 else 0) | _lib.STEM_INDIRECT, self.B, self.H, self.W,
                                      _lib.ptr(wd), _lib.ptr(bd), co, o.ptr(), o.cs, _lib.ACT["silu" if p["act"] else "none"], _lib.ptr(ps0),
                                      desc=f"stem 3->{co} + GAP partial sums")
                        else:
                            self.emit("ysod_stem_mma", _lib.ptr(self.img_slot), (1 if self.src_u8 else 0) | _lib.STEM_INDIRECT, self.B, self.H, self.W, _lib.ptr(wd),
                                      _lib.ptr(bd), co, o.ptr(), o.cs, _lib.ACT["silu" if p["act"] else "none"], desc=f"stem 3->{co}")
                    else:
                        if self.src_u8:
                            raise NotImplementedError("uint8 frame input needs the bf16 tensor-core stem (3x3/s2, Cout 16/32/64)")
                        wd, bd = self.dev_t(wf.permute(0, 2, 3, 1)), self.dev_t(bf)
                        self._stem_direct_op = len(self.ops)
                        self.emit("ysod_stem_conv", None, self.B, self.H, self.W, _lib.ptr(wd), _lib.ptr(bd), co, k, s,
                                  pd, o.ptr(), self.code, o.cs, _lib.ACT["silu" if p["act"] else "none"])
                elif L.i in fold_cv1:
                    J = layers[fold_cv1[L.i]]
                    cj = int(J.c2 * J.p["e"])
                    cat = self.new(self.B, h, w, (2 if self.c2f_fuses_cv2(J.p, J.c2, h, w) else 2 + J.p["n"]) * cj)
                    self._c2f_pre[J.i] = cat
                    self.conv_bn(x, P, p["k"], p["s"], p["g"], p["act"], out=cat.slice(0, 2 * cj), pad=p["p"], gate=gate)
                    w2, b2 = self.folded(f"model.{J.i}.cv1")
                    w2d, b2d = self.dev_t(w2.view(64, 64), self.dt), self.dev_t(b2)
                    self.call("ysod_conv_tc_set_b2b_conv", self.tc_handles[-1], _lib.ptr(w2d), _lib.ptr(b2d), _lib.ACT["silu"])
                    fl2 = 2.0 * self.B * h * w * 64 * 64
                    self.op_flops[-1] += fl2
                    self.tc_flops += fl2
                    self.op_desc[-1] += f" +L{J.i} C2f.cv1 64->64 (b2b)"
                    o = None    # the layer's own output is never materialised (its only consumer is the fused cv1)
                else:
                    self.conv_bn(x, P, p["k"], p["s"], p["g"], p["act"], out=o, pad=p["p"], gate=gate)
            elif t == "C2f":
                self.c2f(x, P, p, L.c2, o)
            elif t == "C3k2":
                self.c2f(x, P, p, L.c2, o, c3k2=True)
            elif t == "A2C2f":
                self.a2c2f(x, P, p, L.c2, o)
            elif t == "SPPF":
                self.sppf(x, P, p, o)
            elif t in ("SE_Block", "SE"):
                self.se(x, P, o)
            elif t == "CBAM_Block":
                self.cbam(x, P, o)
            elif t == "CA_Block":
                self.ca(x, P, o)
            elif t == "SwinBlock":
                self.swin(x, P, p, o)
            elif t == "A2_Attn":
                self.a2attn(x, P, p, o)
            elif t == "MambaBlock":
                self.mamba(x, P, p, o)
            elif t == "nn.Upsample":
                self.emit("ysod_upsample_copy", x.ptr(), self.code, x.N, x.H, x.W, x.C, x.cs, p["scale"], o.ptr(), o.cs)
            else:
                raise NotImplementedError(t)
            out[L.i] = o
            if L.i in det_inputs:
                self.sched.append(("record", f"L{L.i}", 0))

    # ---- execution ---------------------------------------------------------------------------------------
    def launch_all(self):
        """Issues the schedule: lane 0 on the current stream, Detect-level branches on side streams forked / joined with events
        (under graph capture these become graph edges, so independent branches run concurrently inside one CUDA graph)."""
        main = torch.cuda.current_stream()
        if self.n_lanes > 1 and not hasattr(self, "_side"):
            self._side = [torch.cuda.Stream(device=self.dev) for _ in range(self.n_lanes - 1)]
        streams = [main] + (self._side if self.n_lanes > 1 else [])
        ptrs = [C.c_void_p(s.cuda_stream) for s in streams]
        marks = {}
        for kind, what, lane in self.sched:
            if kind == "op":
                fn, args, name = self.ops[what]
                rc = fn(*args, ptrs[lane])
                if rc:
                    _lib.check(rc, name, self.half)
            elif kind == "record":
                ev = torch.cuda.Event()
                ev.record(streams[lane])
                marks[what] = ev
            else:
                streams[lane].wait_event(marks[what])

    def profile(self, iters=3):
        """Eager replay with a CUDA event pair around every launch (on the launching stream). Returns
        {kernel entry point: {"ms": mean ms per forward, "launches": n per forward, "flops": algorithmic flops per forward}}."""
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        acc: Dict[str, list] = {}
        per_op = [0.0] * len(self.ops)
        for it in range(iters + 1):
            evs = []
            for fn, args, name in self.ops:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = fn(*args, st)
                e1.record()
                if rc:
                    _lib.check(rc, name, self.half)
                evs.append((e0, e1))
            torch.cuda.synchronize()
            if it == 0:
                continue  # warm-up pass
            for i, ((fn, args, name), (e0, e1), fl) in enumerate(zip(self.ops, evs, self.op_flops)):
                a = acc.setdefault(name, [0.0, 0, 0.0])
                per_op[i] += e0.elapsed_time(e1)
                a[0] += e0.elapsed_time(e1)
                a[1] += 1
                a[2] += fl
        self.last_per_op = [{"op": i, "kernel": self.ops[i][2], "desc": self.op_desc[i], "ms": per_op[i] / iters,
                             "gflop": self.op_flops[i] / 1e9} for i in range(len(self.ops))]
        return {k: {"ms": v[0] / iters, "launches": v[1] // iters, "flops": v[2] / iters} for k, v in acc.items()}

    def capture(self):
        """Capture the launch list into a CUDA graph (batch-1 latency is launch-bound otherwise)."""
        warm = None
        if self.img_indirect and not self._bound:
            warm = torch.zeros(self.img_shape, device=self.dev, dtype=self.img_dtype)   # warm-up image; every run() rebinds
            self.bind_input(warm)
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.launch_all()  # warm-up outside capture
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.launch_all()
        self.graph = g
        if warm is not None:
            torch.cuda.synchronize()
            self._bound = 0
            del warm

    def _rebind_direct(self):
        """CUDA-core stem (fp32 parity mode): the kernel takes the image pointer as a plain argument = the private staging buffer."""
        i = self._stem_direct_op
        fn, args, name = self.ops[i]
        self.ops[i] = (fn, (_lib.ptr(self.img),) + tuple(args[1:]), name)

    def bind_input(self, x: torch.Tensor):
        """Points the stem at `x` (no copy) when it is a dense, 16 B aligned tensor of the program's dtype / layout on its device;
        anything else (host memory, other dtypes, strided views) is staged through the private buffer."""
        direct = (self.img_indirect and x.is_cuda and x.device == self.dev and x.dtype == self.img_dtype
                  and tuple(x.shape) == self.img_shape and x.is_contiguous() and x.data_ptr() % 16 == 0)
        if not direct:
            if self.img is None:
                self.img = torch.zeros(self.img_shape, device=self.dev, dtype=self.img_dtype)
            self.img.copy_(x, non_blocking=True)
            x = self.img
        if self.img_indirect and x.data_ptr() != self._bound:
            self.call("ysod_set_ptr", _lib.ptr(self.img_slot), C.c_void_p(x.data_ptr()), _lib.stream_ptr())
            self._bound = x.data_ptr()

    def run(self, x: torch.Tensor, static: bool = False):
        self.bind_input(x)
        if self.graph is not None:
            self.graph.replay()
        else:
            self.launch_all()
        raw = [r.torch_nchw() for r in self.raws]
        if static:
            return self.y, raw      # views of the program's own buffers: overwritten by the next forward of this shape
        return self.y.clone(), [r.clone() for r in raw]

    def __del__(self):
        try:
            for h in self.tc_handles:
                self.lib.ysod_conv_tc_destroy(h)
        except Exception:
            pass


class B200DetectionModel:
    """Drop-in for the reference's `DetectionModel` forward seam (tasks.py:129-163): `model(x) -> (y, [raw...])`.

    cfg        : config name / YAML path / dict (see cfg.py)
    state_dict : reference-named weights (e.g. `ref_model.state_dict()`); fp32 CPU or CUDA tensors
    dtype      : torch.bfloat16 (tensor-core path), torch.float16 (the same path on the fp16 build of the library: the reference's
                 `half=True` / `model.half()` mode, autobackend.py:154) or torch.float32 (CUDA-core parity mode, rtol 1e-4)
    """

    def __init__(self, cfg, state_dict, dtype=torch.bfloat16, device="cuda:0", use_tc=True, use_graph=True, nc=None,
                 static_outputs=False, max_programs=8, attn_impl=0, fuse_cbam=False, ca_single_pass=False, fuse_b2b=True, swin_nhwc=True, fuse_swin=True, multi_stream=True, fuse_upsample=True, fuse_decode=True, fuse_se=True, fuse_gate=False, conv_duo=True, swin_impl=0, fuse_stem_gap=True):
        _lib.require_cuda()
        _lib.load()
        self.spec = cfg if isinstance(cfg, _cfg.ModelSpec) else _cfg.get_spec(cfg, nc=nc)
        self.sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        missing = [k for k in _cfg.param_shapes(self.spec) if k not in self.sd]
        if missing:
            raise KeyError(f"state_dict is missing {len(missing)} tensors, e.g. {missing[:3]}")
        assert dtype in (torch.bfloat16, torch.float16, torch.float32)
        self.dtype = dtype
        _lib.load(dtype == torch.float16)
        self.device = torch.device(device)
        self.use_tc = use_tc
        self.use_graph = use_graph
        self.fuse_swin = fuse_swin
        self.swin_impl = int(os.environ.get("YSOD_SWIN_IMPL", swin_impl))   # fused P2 SwinBlock: 0 = tcgen05 kernel, 1 = mma.sync kernel (A/B)
        if "YSOD_FUSE_SWIN" in os.environ:
            self.fuse_swin = os.environ["YSOD_FUSE_SWIN"] == "1"
        # Single-pass variants of the CBAM spatial stage (ysod_cbam_spatial) and of the CoordAtt pooling (ysod_ca_pool with a workspace):
        # bit-identical / fp32-rounding-identical to the multi-pass kernels but measured SLOWER on B200 (profiles/r02_ab_blocks.json:
        # 128 vs 107 us at 64 x 160^2, 92 vs 42 us at 256 x 40^2; 120 vs 37 us for the pooling) -- one CTA per tile serialises statistics ->
        # filter -> apply with too few bytes in flight, while the separate passes stream at full occupancy. Off by default.
        self.fuse_cbam = fuse_cbam
        self.ca_single_pass = ca_single_pass
        self.fuse_b2b = fuse_b2b     # Detect branch 3x3 conv -> 1x1 head conv -> decode in one launch (ysod_conv_tc_set_b2b)
        self.swin_nhwc = swin_nhwc   # unfused SwinBlocks keep their tokens in NHWC order (no window_partition / window_reverse copies)
        if "YSOD_SWIN_NHWC" in os.environ:
            self.swin_nhwc = os.environ["YSOD_SWIN_NHWC"] == "1"
        if "YSOD_FUSE_B2B" in os.environ:       # A/B switches for measurements
            self.fuse_b2b = os.environ["YSOD_FUSE_B2B"] == "1"
        if "YSOD_MULTI_STREAM" in os.environ:
            multi_stream = os.environ["YSOD_MULTI_STREAM"] == "1"
        self.conv_pair = os.environ.get("YSOD_NO_PAIR", "0") != "1"   # A/B switch of the conv kernel's tile-pair plan (ysod.h YSOD_CONV_NO_PAIR)
        # pixel-duo plan for dense 32 -> 32 3x3 convs (ysod.h YSOD_CONV_NO_DUO): same products, different accumulation order than the 32-channel plan
        self.fuse_stem_gap = fuse_stem_gap and os.environ.get("YSOD_STEM_GAP", "1") != "0"   # stem kernel leaves the SE pooling partials (A/B switch)
        self.conv_duo = conv_duo and os.environ.get("YSOD_NO_DUO", "0") != "1"
        self.c2f_cat = os.environ.get("YSOD_C2F_CAT", "1") == "1"     # A/B switch: C2f.cv2 inside the last Bottleneck conv's launch (ysod_conv_tc_set_b2b_cat)
        self.attn_impl = attn_impl   # ysod_mha_core_ex impl: 0 = tcgen05 / TMEM attention core where covered, 1 = mma.sync kernels (A/B)
        self.fuse_upsample = fuse_upsample
        self.fuse_decode = fuse_decode
        self.fuse_se = fuse_se
        self.fuse_gate = fuse_gate   # pool + gate MLP in one launch (ysod_gap_gate): measured 0.5 % slower at B=32 and 8 us slower at B=1
                                     # than the separate PDL-chained gate launch (the last CTA's serial tail), so off by default
        self.multi_stream = multi_stream
        self.static_outputs = static_outputs
        self.stride_list = _cfg.strides_of(self.spec)
        self.stride = torch.tensor([float(s) for s in self.stride_list])
        self.nc = self.spec.nc
        self.names = {i: f"{i}" for i in range(self.nc)}  # tasks.py:351 default names
        self.yaml = self.spec.yaml
        self.reg_max = 16
        self.no = self.nc + 64
        self.nl = len(self.stride_list)
        # One compiled program (activation buffers + CUDA graph) per input signature, least recently used evicted beyond
        # `max_programs`: rectangular letterbox shapes (LetterBox auto=True) would otherwise grow GPU memory without bound.
        self.max_programs = max(1, int(max_programs))
        self.programs: "OrderedDict[tuple, Program]" = OrderedDict()

    def program(self, B, H, W, src_u8=False, want_raw=True, slot=0) -> Program:
        """slot: independent copies of a program (own activation buffers, `y` and CUDA graph) for callers that keep two forwards in
        flight, e.g. YOLO's overlap of step i's NMS with step i+1's forward."""
        key = (B, H, W, bool(src_u8), bool(want_raw), int(slot))
        if key in self.programs:
            self.programs.move_to_end(key)
        else:
            s = max(self.stride_list)
            if H % s or W % s:
                raise ValueError(f"input {H}x{W} must be a multiple of the max stride {s}")
            with torch.cuda.device(self.device):
                prog = Program(self, B, H, W, src_u8, want_raw)
                if self.use_graph:
                    prog.capture()
            self.programs[key] = prog
            while len(self.programs) > self.max_programs:
                self.programs.popitem(last=False)   # buffers / graph / conv plans are released with the Program
        return self.programs[key]

    @torch.no_grad()
    def forward(self, x, *args, static=None, want_raw=True, slot=0, **kwargs):
        """x: (B,3,H,W) float tensor in [0,1] (NCHW, as the reference's forward takes it), or (B,H,W,3) uint8 BGR frames (what
        BasePredictor.preprocess receives, predictor.py:116-134; BGR->RGB, HWC->CHW and /255 are fused into the stem kernel).
        Returns (y, [raw maps]) as fresh tensors, like the reference's forward. `static=True` (or `static_outputs=True` at
        construction) returns views of the program's own buffers instead -- zero-copy, but overwritten by the next forward of the
        same shape: for callers that consume the result immediately (YOLO.predict, bench). `want_raw=False` (predict path: only
        `y` is consumed) skips materialising the raw maps where the decode is fused into the head conv; the list is then empty."""
        u8 = x.dtype == torch.uint8
        if x.dim() != 4 or (x.shape[3] if u8 else x.shape[1]) != 3:
            raise ValueError(f"expected (B,3,H,W) float or (B,H,W,3) uint8, got {tuple(x.shape)} {x.dtype}")
        # host tensors are copied straight into the program's input buffer (pinned memory makes this asynchronous)
        if u8:
            prog = self.program(int(x.shape[0]), int(x.shape[1]), int(x.shape[2]), True, want_raw, slot)
        else:
            prog = self.program(int(x.shape[0]), int(x.shape[2]), int(x.shape[3]), False, want_raw, slot)
        with torch.cuda.device(self.device):
            return prog.run(x, self.static_outputs if static is None else bool(static))

    __call__ = forward
    predict = forward

    def eval(self):
        return self

    def half(self):
        """`model.half()` (autobackend.py:154): the same graph and weights compiled against the fp16 build of the library. Returns a
        new model object (compiled programs are per dtype); `self` is left untouched."""
        if self.dtype == torch.float16:
            return self
        import copy
        twin = copy.copy(self)
        twin.dtype = torch.float16
        twin.programs = OrderedDict()
        _lib.load(True)
        return twin

    def fuse(self, verbose=False):
        return self  # BN is always folded at compile time

    def layer_output(self, x, idx):
        """Debug/test helper: NCHW fp32 copy of layer `idx`'s output for the last forward with this shape."""
        u8 = x.dtype == torch.uint8
        prog = self.program(int(x.shape[0]), int(x.shape[1 if u8 else 2]), int(x.shape[2 if u8 else 3]), u8)
        v = prog.layer_out[idx]
        return None if v is None else v.torch_nchw().float()   # None: fused away (e.g. a Conv whose only consumer runs inside its launch)
