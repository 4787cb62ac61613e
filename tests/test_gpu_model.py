"""GPU parity of the whole forward path (engine + every kernel) against the oracle and the live-reference goldens.

Tolerances (north_star: raw head maps rtol 1e-4 in fp32 mode, rtol 2e-2 in bf16 mode):
  fp32 mode : elementwise |got-ref| <= 1e-4*|ref| + 1e-4*max|ref|   (absolute floor for near-zero elements)
  bf16 mode : relative L2 error ||got-ref|| / ||ref|| <= 2e-2 on every raw head map (3e-2 on intermediate layers, which
              are diagnostics), at most 0.3 % of the elements outside |got-ref| <= 2e-2*|ref| + 2e-2*max|ref|, and a hard cap
              of 8 % of max|ref| on any single element (both tightened in round 2 from the recorded statistics).
  fp16 mode : the reference's own `half=True`; rel-L2 <= 3e-3 against the fp32 oracle and <= 4e-3 against the oracle itself run
              in fp16 (the reference's fp16 path), strict elementwise bound at rtol 1e-2.
A pure elementwise rtol 2e-2 in the max norm is not what a bf16 execution of this ~100-layer network delivers: the
reference's OWN code run in bf16 (model.fuse().bfloat16() on CPU) deviates from its fp32 run by 3.1-5.1 % of max|ref|
(SOD, 160^2 and 640^2) and 4.5-9.9 % (yolov12n 640^2) on single elements, rel-L2 1.0-2.2 % -- measured in the build
container, numbers in DESIGN.md. This path measures 2.3-5.4 % on the same inputs."""
import json
import os

import numpy as np
import pytest
import torch

import yolo_sod_b200  # noqa: F401
from yolo_sod_b200 import cfg as ycfg, synth
from oracle import model_ref, nms_ref
from tests.golden.make_golden_model import CASES as GOLD_CASES

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "model_golden.npz")
SOD = "yolov12-sod-fusion-v5-simple"


# bf16 allowances, set from the recorded statistics (profiles/r02_parity_records.jsonl: 78 raw maps over every model test):
# measured share of elements outside rtol*|ref| + rtol*max|ref| is 0 for 68 maps, <= 1.1e-4 for the SOD family, 1.8e-3 worst case
# (yolov12n P5 map at 640^2); measured max single-element error 5.9 % of max|ref| (same map), <= 3.6 % for the SOD family.
VIOL_SHARE = 3e-3     # allowed share of elements outside the strict bound (round 1: 1e-2)
MAX_ERR_CAP = 0.08    # hard cap on any single element, as a fraction of max|ref| (round 1: 0.10)
PARITY_LOG = os.environ.get("YSOD_PARITY_LOG", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out",
                                                            "parity_records.jsonl"))


def _record(**kw):
    """Appends one parity measurement (rel-L2, max err / max|ref|, share of elements outside the strict bound) per raw map and test
    to a JSON-lines file, so the bf16 allowance can be tracked and tightened (copied to profiles/ per round)."""
    try:
        os.makedirs(os.path.dirname(PARITY_LOG), exist_ok=True)
        with open(PARITY_LOG, "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass


def _stats(got, ref, tol):
    err = (got - ref).abs()
    mx = ref.abs().max()
    return {"rel_l2": float(err.norm() / ref.norm().clamp_min(1e-12)), "max_err": float(err.max()), "max_ref": float(mx),
            "max_err_over_max_ref": float(err.max() / mx.clamp_min(1e-30)),
            "viol_share": float((err > tol * ref.abs() + tol * mx).float().mean())}


def _close(got, ref, tol):
    """tol < 1e-2: fp32 criterion; else the bf16 criterion (see module docstring). Returns (ok, max err, max|ref|, rel-L2)."""
    st = _stats(got, ref, tol)
    if tol < 1e-2:
        return st["viol_share"] == 0.0, st["max_err"], st["max_ref"], st["rel_l2"]
    ok = st["rel_l2"] <= tol and st["viol_share"] <= VIOL_SHARE and st["max_err_over_max_ref"] <= MAX_ERR_CAP
    return ok, st["max_err"], st["max_ref"], st["rel_l2"]


def _build(name, dtype, **kw):
    from yolo_sod_b200.model import DetectionModel
    spec = ycfg.get_spec(name)
    sd = synth.synth_state_dict(spec, name, 0)
    return spec, sd, DetectionModel(name, weights=sd, dtype=dtype, **kw)


def _check(name, dtype, B, sz, tol, layer_tol=None, seed=11, layers_too=True, tag=None, **kw):
    """layers_too=False: the full-size configs compare the raw head maps and `y` only (keeping every layer of a 32-image 640^2
    batch in fp32 on the host is ~7 GB); per-layer parity is covered by the small cases."""
    spec, sd, model = _build(name, dtype, **kw)
    x = synth.synth_images(B, sz, seed=seed)
    y, raw = model(x.cuda())
    torch.cuda.synchronize()
    if layers_too:
        (y_ref, raw_ref), layers = model_ref.forward(spec, sd, x, ycfg.strides_of(spec), return_layers=True)
    else:
        y_ref, raw_ref = model_ref.forward(spec, sd, x, ycfg.strides_of(spec))
        layers = None
    report = []
    for i in range(len(spec.layers) - 1 if layers_too else 0):
        got = model.layer_output(x, i)
        if got is None:          # not materialised: the layer's only consumer runs inside its launch (back-to-back GEMM)
            continue
        got = got.cpu()
        ok, e, m, l2 = _close(got, layers[i], layer_tol or tol)
        report.append((i, spec.layers[i].type, ok, round(e, 5), round(m, 3), round(l2, 4)))
    bad = [r for r in report if not r[2]]
    caller = tag or os.environ.get("PYTEST_CURRENT_TEST", "").split("::")[-1].split(" ")[0]
    for l, (a, b) in enumerate(zip(raw, raw_ref)):
        st = _stats(a.float().cpu(), b, tol)
        _record(test=caller, model=name, dtype=str(dtype).replace("torch.", ""), batch=B, imgsz=sz, raw_map=l, tol=tol, **st)
        ok, e, m, l2 = _close(a.float().cpu(), b, tol)
        assert ok, f"raw map {l}: max err {e} (ref max {m}) rel-L2 {l2} viol {st['viol_share']}; first bad layers: {bad[:4]}"
    assert not bad, f"layer mismatches: {bad[:6]}"
    # decoded boxes / scores: fp32 decode of our own raw maps must match the oracle decode of the same maps
    y_dec = model_ref.detect_decode([r.float().cpu() for r in raw], ycfg.strides_of(spec), spec.nc)
    assert torch.allclose(y.cpu(), y_dec, rtol=1e-4, atol=2e-3), float((y.cpu() - y_dec).abs().max())
    return model, x, y, y_ref


def test_sod_fp32_mode():
    _check(SOD, torch.float32, 2, 128, 1e-4, layer_tol=2e-4)


def test_sod_bf16_tensor_core_path():
    model, x, y, y_ref = _check(SOD, torch.bfloat16, 2, 160, 2e-2, layer_tol=3e-2)
    prog = model.program(2, 160, 160)
    assert prog.n_tc > 60, "the dense convs must run on the tcgen05 kernel"


def test_sod_fp16_half_mode():
    """the reference's only native low-precision mode, `half=True` -> model.half() (autobackend.py:154). The fp16
    build of the library (same sources, -DYSOD_HALF=1: IEEE fp16 storage and tensor-core inputs) against (a) the fp32 oracle and
    (b) the oracle run in fp16 itself = the reference's fp16 path."""
    spec, sd, model = _build(SOD, torch.float16)
    x = synth.synth_images(2, 160, seed=11)
    y, raw = model(x.cuda())
    torch.cuda.synchronize()
    assert model.program(2, 160, 160).n_tc > 60, "the dense convs must run on the tcgen05 kernel"
    y32, raw32 = model_ref.forward(spec, sd, x, ycfg.strides_of(spec))
    fsd = {k: (v.half() if v.is_floating_point() else v) for k, v in model_ref.fuse_state_dict(sd).items()}
    y16, raw16 = model_ref.forward(spec, fsd, x, ycfg.strides_of(spec), dtype=torch.float16)
    for l, (a, b32, b16) in enumerate(zip(raw, raw32, raw16)):
        st = _stats(a.float().cpu(), b32, 1e-2)
        _record(test="test_sod_fp16_half_mode", model=SOD, dtype="float16", batch=2, imgsz=160, raw_map=l, tol=1e-2, **st)
        assert st["rel_l2"] <= 3e-3 and st["viol_share"] == 0.0, (l, st)
        st16 = _stats(a.float().cpu(), b16.float(), 1e-2)
        assert st16["rel_l2"] <= 4e-3, (l, st16)
    # the public switches: YOLO.predict(half=True) and AutoBackend(fp16=True) run the fp16 twin of a bf16 model
    from yolo_sod_b200.model import YOLO, AutoBackend
    yolo = YOLO(SOD, weights=sd, dtype=torch.bfloat16)
    be = AutoBackend(yolo.model, fp16=True)
    assert be.fp16 and torch.equal(be(x.cuda())[0], y)
    r16 = yolo.predict(x.cuda(), half=True)
    r_bf = yolo.predict(x.cuda())
    assert len(r16) == len(r_bf) == 2 and sum(len(r) for r in r16) > 0
    # bf16 and fp16 libraries coexist in one process
    _check(SOD, torch.bfloat16, 2, 160, 2e-2, layer_tol=3e-2)


def test_sod_bf16_cuda_core_crosscheck():
    _check(SOD, torch.bfloat16, 1, 128, 2e-2, layer_tol=3e-2, use_tc=False)


def test_sod_bf16_640_batch2():
    _check(SOD, torch.bfloat16, 2, 640, 2e-2, layer_tol=3e-2, seed=99)


def test_sod_bf16_c2_stated_config_640_batch32():
    """BASELINE config C2 at its stated size: SOD-simple, 640^2, batch 32, bf16 -- raw head maps and y of all 32 images vs the oracle."""
    _check(SOD, torch.bfloat16, 32, 640, 2e-2, seed=41, layers_too=False)


def test_sod_bf16_c3_stated_config_1024_batch16():
    """BASELINE config C3 at its stated size: SOD-simple, 1024^2, batch 16, bf16."""
    model, x, y, y_ref = _check(SOD, torch.bfloat16, 16, 1024, 2e-2, seed=43, layers_too=False)
    assert y.shape == (16, 14, 87040)


def test_yolov12m_c4b_640_bf16_and_fp32_mode():
    """BASELINE config C4 read as stock yolov12m (SURVEY 8d option ii) at 640^2: bf16 tensor-core path on a micro-batch, and the fp32
    parity mode (rtol 1e-4) of the same model."""
    _check("yolov12m", torch.bfloat16, 4, 640, 2e-2, seed=45, layers_too=False)
    _check("yolov12m", torch.float32, 1, 128, 1e-4, layer_tol=2e-4, seed=45)


def test_sod_bf16_1024_single_image():
    """BASELINE config C3 (VisDrone-scale 1024^2: 87 040 anchors, 1369 P2 windows of which the last row / column are zero-padded)."""
    model, x, y, y_ref = _check(SOD, torch.bfloat16, 1, 1024, 2e-2, layer_tol=3e-2, seed=5)
    assert y.shape[2] == 87040


@pytest.mark.parametrize("name", ["E1", "E2", "E3", "E4", "E5", "E6"])
def test_ablation_ladder_bf16(name):
    """The paper's ablation models (cfg/models/new/E1..E6.yaml; E1 has three heads, no P2): same kernels, fewer blocks."""
    _check(name, torch.bfloat16, 1, 160, 2e-2, layer_tol=4e-2, seed=17)   # raw head maps at 2 %; intermediate layers are diagnostics (4 %)


def test_stable_variant_bf16():
    """cfg/models/new/yolov12-sod-fusion-v5-stable.yaml: DetectStable head (== Detect in eval), Swin blocks at P5 (8 heads, 512 ch) and
    P2 (4 heads: the unfused window-attention path, since the fused kernel is the 2-head instance)."""
    _check("yolov12-sod-fusion-v5-stable", torch.bfloat16, 1, 320, 2e-2, layer_tol=4e-2, seed=19)


def test_full_v5_with_mamba_glu_fallback():
    """cfg/models/new/yolov12-sod-fusion-v5.yaml: the -simple graph + MambaBlock, which the reference runs through its GLU fallback when
    mamba_ssm is not installed (blocks_mamba.py:116-165) -- fp32 parity mode and the bf16 tensor-core path."""
    _check("yolov12-sod-fusion-v5", torch.float32, 1, 128, 1e-4, layer_tol=2e-4, seed=21)
    _check("yolov12-sod-fusion-v5", torch.bfloat16, 1, 320, 2e-2, layer_tol=4e-2, seed=21)


def test_yolov12n_fp32_and_bf16():
    _check("yolov12n", torch.float32, 2, 128, 1e-4, layer_tol=2e-4)
    _check("yolov12n", torch.bfloat16, 1, 640, 2e-2, layer_tol=4e-2)  # attention-heavy: diagnostic layer bound 4 %


def test_yolov12m_bf16():
    _check("yolov12m", torch.bfloat16, 1, 128, 2e-2, layer_tol=3e-2)


@pytest.mark.parametrize("case", GOLD_CASES, ids=lambda c: c[0])
def test_against_live_reference_goldens(case):
    key, name, _, b, sz = case
    g = np.load(GOLD)
    spec, sd, model = _build(name, torch.float32)
    x = synth.synth_images(b, sz, seed=7)
    y, raw = model(x.cuda())
    assert np.allclose(y.cpu().numpy(), g[f"{key}_y"], rtol=1e-3, atol=5e-3)
    for l, r in enumerate(raw):
        ok, e, m, l2 = _close(r.float().cpu(), torch.from_numpy(g[f"{key}_raw{l}"].astype(np.float32)), 2e-3)
        assert ok, (key, l, e, m, l2)
    spec, sd, model = _build(name, torch.bfloat16)
    y, raw = model(x.cuda())
    for l, r in enumerate(raw):
        ok, e, m, l2 = _close(r.float().cpu(), torch.from_numpy(g[f"{key}_raw{l}"].astype(np.float32)), 2e-2)
        assert ok, (key, l, e, m, l2)


def test_graph_replay_equals_eager_and_is_deterministic():
    spec, sd, m_graph = _build(SOD, torch.bfloat16, use_graph=True)
    _, _, m_eager = _build(SOD, torch.bfloat16, use_graph=False)
    x = synth.synth_images(2, 128, seed=3).cuda()
    y1 = m_graph(x)[0].clone()
    y2 = m_eager(x)[0].clone()
    y3 = m_graph(x)[0].clone()
    assert torch.equal(y1, y2) and torch.equal(y1, y3)
    # the Detect-level branches on side streams (multi_stream) must not change a single bit either
    _, _, m_single = _build(SOD, torch.bfloat16, multi_stream=False)
    assert m_graph.program(2, 128, 128).n_lanes == 4 and m_single.program(2, 128, 128).n_lanes == 1
    for _ in range(3):
        assert torch.equal(m_graph(x)[0], m_single(x)[0])


def test_outputs_are_fresh_tensors_and_predict_path_skips_raw_maps():
    """Drop-in semantics: forward returns fresh tensors (a second forward must not overwrite the first result), the
    static / no-raw-map mode used by YOLO.predict gives the bit-identical `y`, strided or host inputs are staged, and the program
    cache is bounded."""
    spec, sd, model = _build(SOD, torch.bfloat16, max_programs=2)
    xa, xb = synth.synth_images(2, 160, seed=1).cuda(), synth.synth_images(2, 160, seed=2).cuda()
    ya, ra = model(xa)
    keep = ya.clone()
    yb, rb = model(xb)
    assert torch.equal(ya, keep) and not torch.equal(ya, yb), "results must not alias the program's buffers"
    ys, rs = model(xa, static=True, want_raw=False)
    assert rs == [] and torch.equal(ys, ya), "decode-only (no raw map) program must give the same y"
    # the same values through the staging path: a host tensor, and a strided (non-contiguous) device view
    yh, _ = model(xa.cpu())
    big = torch.zeros((2, 3, 160, 320), device="cuda")
    big[..., ::2] = xa
    yv, _ = model(big[..., ::2])
    assert torch.equal(yh, ya) and torch.equal(yv, ya)
    for sz in (128, 192, 224):
        model(synth.synth_images(1, sz, seed=3).cuda())
    assert len(model.programs) == 2, "LRU cap on compiled programs"


def test_predict_end_to_end_matches_oracle_nms():
    from yolo_sod_b200.model import YOLO
    from yolo_sod_b200 import ops
    yolo = YOLO(SOD, dtype=torch.bfloat16)
    x = synth.synth_images(2, 320, seed=21)
    res = yolo.predict(x.cuda(), conf=0.25, iou=0.7, max_det=300)
    y, _ = yolo.model(x.cuda())
    want = nms_ref.non_max_suppression(y.cpu().numpy(), 0.25, 0.7, max_det=300)
    assert len(res) == 2
    for r, w in zip(res, want):
        w = w.copy()
        w[:, [0, 2]] = w[:, [0, 2]].clip(0, 320)
        w[:, [1, 3]] = w[:, [1, 3]].clip(0, 320)
        assert np.array_equal(r.boxes.data.cpu().numpy(), w)
    assert sum(len(r) for r in res) > 0, "synthetic weights must produce detections"
    # stream=True (engine/model.py:501-560): a generator over the same Results; consuming it after a later batch was submitted
    # must still give this batch's detections (at most 4 unconsumed generators may be outstanding: the count ring)
    g1 = yolo.predict(x.cuda(), stream=True, conf=0.25, iou=0.7, max_det=300)
    g2 = yolo.predict(synth.synth_images(2, 320, seed=22).cuda(), stream=True)
    for r, w in zip(list(g1), res):
        assert torch.equal(r.boxes.data, w.boxes.data)
    assert len(list(g2)) == 2


def test_overlap_nms_mode_gives_identical_results():
    """YOLO(overlap_nms=True): batch i's NMS on a side stream under batch i+1's forward, two program slots. Same detections as the
    serial mode for a sequence of different batches, through predict(), predict(stream=True) and predict_padded() + join()."""
    from yolo_sod_b200.model import YOLO
    spec = ycfg.get_spec(SOD)
    sd = synth.synth_state_dict(spec, SOD, 0)
    serial = YOLO(SOD, weights=sd, dtype=torch.bfloat16)
    piped = YOLO(SOD, weights=sd, dtype=torch.bfloat16, overlap_nms=True)
    batches = [synth.synth_images(2, 320, seed=30 + i).cuda() for i in range(5)]
    want = [[r.boxes.data.clone() for r in serial.predict(b)] for b in batches]
    gens = [piped.predict(b, stream=True) for b in batches[:4]]          # four batches in flight before the first is consumed
    for w, g in zip(want, gens):
        for a, r in zip(w, list(g)):
            assert torch.equal(a, r.boxes.data)
    for w, b in zip(want, batches):
        for a, r in zip(w, piped.predict(b)):
            assert torch.equal(a, r.boxes.data)
    outs = [piped.predict_padded(b) for b in batches]
    piped.join()
    torch.cuda.synchronize()
    for w, (det, cnt, _) in zip(want, outs):
        for b_, a in enumerate(w):
            assert int(cnt[b_]) == a.shape[0] and torch.equal(det[b_, :a.shape[0]], a)


def test_fused_swin_block_matches_oracle_and_unfused_path():
    """Layer 28 (P2 SwinBlock, 64 ch): the single fused kernel vs the oracle layer output and vs the eleven-launch path.
    161 is not a multiple of 7 at 640^2 -> 160x160 maps exercise the zero-padded windows; 96^2 input -> 24x24 map (4x4 windows, ragged)."""
    for sz in (96, 224):
        spec, sd, m_f = _build(SOD, torch.bfloat16)
        _, _, m_u = _build(SOD, torch.bfloat16, fuse_swin=False)
        x = synth.synth_images(2, sz, seed=13)
        m_f(x.cuda()); m_u(x.cuda())
        torch.cuda.synchronize()
        (_, _), layers = model_ref.forward(spec, sd, x, ycfg.strides_of(spec), return_layers=True)
        got_f, got_u, ref = m_f.layer_output(x, 28).cpu(), m_u.layer_output(x, 28).cpu(), layers[28]
        ok, e, m, l2 = _close(got_f, ref, 3e-2)
        assert ok, ("fused vs oracle", sz, e, m, l2)
        ok, e, m, l2u = _close(got_u, ref, 3e-2)
        assert l2 <= l2u * 1.25 + 1e-3, ("fused path must not be less accurate than the unfused one", l2, l2u)
        assert any(o[2] == "ysod_swin64_tc" for o in m_f.program(2, sz, sz).ops)
        # the mma.sync version of the fused kernel (A/B baseline): same rounding points, so the two agree to accumulation order
        _, _, m_s = _build(SOD, torch.bfloat16, swin_impl=1)
        m_s(x.cuda())
        torch.cuda.synchronize()
        assert any(o[2] == "ysod_swin64_fused" for o in m_s.program(2, sz, sz).ops)
        got_s = m_s.layer_output(x, 28).cpu()
        ok, e, m, l2s = _close(got_s, ref, 3e-2)
        assert ok, ("mma.sync fused vs oracle", sz, e, m, l2s)
        d = (got_f.float() - got_s.float()).abs()
        assert float(d.max()) <= 0.05 * float(ref.abs().max()) and float((d > 0.01 * float(ref.abs().max())).float().mean()) < 1e-3, float(d.max())


def test_fused_epilogues_equal_separate_kernels():
    """Conv+Upsample in one launch and the Detect decode inside the final head conv's epilogue are pure re-schedulings: the raw maps
    are bit-identical to the separate-kernel program and y agrees to fp32 rounding (expression order is the same; only FMA contraction may differ)."""
    spec, sd, fused = _build(SOD, torch.bfloat16, fuse_gate=True, fuse_cbam=True)     # + pool and gate MLP in one launch (ysod_gap_gate), single-pass CBAM spatial stage
    _, _, plain = _build(SOD, torch.bfloat16, fuse_upsample=False, fuse_decode=False, fuse_gate=False, fuse_cbam=False, fuse_b2b=False,
                         conv_duo=False)   # + CBAM as stats / apply passes; (the pixel-duo conv plan sums its taps in another order: not a pure re-scheduling)
    x = synth.synth_images(2, 320, seed=23).cuda()
    y1, r1 = fused(x)
    y2, r2 = plain(x)
    torch.cuda.synchronize()
    assert fused.program(2, 320, 320).n_launches < plain.program(2, 320, 320).n_launches - 6
    for a, b in zip(r1, r2):
        assert torch.equal(a, b)
    assert torch.allclose(y1, y2, rtol=1e-5, atol=1e-4), float((y1 - y2).abs().max())


def test_uint8_frames_equal_preprocessed_tensor():
    """(B,H,W,3) uint8 BGR frames through the fused stem == the reference's preprocess (predictor.py:127-133) + tensor input."""
    spec, sd, model = _build(SOD, torch.bfloat16)
    g = torch.Generator().manual_seed(5)
    frames = torch.randint(0, 256, (2, 160, 160, 3), generator=g, dtype=torch.uint8)
    x = frames.flip(-1).permute(0, 3, 1, 2).float() / 255
    y_u8 = model(frames.cuda())[0].clone()
    y_f = model(x.cuda())[0].clone()
    assert torch.equal(y_u8, y_f)


def test_missing_library_or_cpu_input_fails_loudly():
    from yolo_sod_b200 import ops, lib
    with pytest.raises(lib.YsodError):
        ops.non_max_suppression(torch.zeros(1, 14, 100), 0.25, 0.7)
