#!/usr/bin/env python
"""Benchmark of the hot path: YOLOv12-SOD forward + decode + NMS, images/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C1|C2|C3|C4|C4b|C5u|C5c]

A "step" is one pass of the hot path (forward graph + decode + batched NMS + box clipping, i.e. `YOLO.predict`'s device work)
over one batch of synthetic images. The default workload is BASELINE.json configs[1] (C2): SOD fusion v5-simple, 640^2, batch 32
per GPU, bf16; N>1 = one process per GPU (torchrun), weak scaling, the only data-path exchange is ONE NCCL all_gather of the
padded detections per step (issued on a side stream so it overlaps the next step's forward). Rank 0 prints one JSON line.

`--config` selects the other BASELINE.json configs (same JSON contract, clock sampler included):
    C1   yolov12n 640^2 batch 1                  C3   SOD 1024^2 batch 16
    C4   SOD 640^2, 256 images sharded over the N GPUs (strong scaling; BASELINE configs[3] read literally, SURVEY 8d option i)
    C4b  yolov12m 640^2, 256 images sharded (SURVEY 8d option ii)
    C5u / C5c  NMS only, 30 000 boxes x 10 classes, batch 64 (sharded), uniform / clustered boxes

`--impl reference` times the reference's own CPU implementation of the same path on the host cores. The reference is pure
Python on torch CPU ops and cannot be installed on the GPU box (its package is incomplete), so the arm
runs the oracle port (oracle/model_ref.py + oracle/nms_ref.py: the same ATen CPU kernels, pinned bit-close to the live
reference) on a bounded sample of the same workload.

Baselines printed next to the B200 numbers (N=1, rank 0): `cpu_baseline` (all host cores) and `cpu_baseline_t8` (the reference's
own thread cap, utils/__init__.py:43), and `gpu_library_baseline`: the same oracle port run on the same B200 in bf16 channels_last
with BN folded (BaseModel.fuse) + torchvision.ops.nms CUDA -- i.e. what the reference's stock GPU path (cuDNN / cuBLAS / ATen)
does for this workload; SURVEY 2.2's bar for the hand-written kernels.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line. Libraries write there too (NCCL prints its "NCCL version ..." banner on stdout), so file
# descriptor 1 is pointed at stderr for the whole run and the result line is written to the saved original stdout by emit().
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
_RESULT_FD = None


def _capture_stdout():
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


SOD = "yolov12-sod-fusion-v5-simple"
CONF, IOU, MAX_DET = 0.25, 0.7, 300
MICRO = 32   # images per forward for the sharded 256-image configs
# kind "model": forward + NMS; "nms": NMS only. total: images per step over all GPUs (strong scaling) or None (weak: batch per GPU)
CONFIGS = {
    "C1": dict(kind="model", model="yolov12n", imgsz=640, batch=1, total=None,
               metric="images/sec (fwd+NMS, 640^2, bf16), batch 1", cpu_batch=1),
    "C2": dict(kind="model", model=SOD, imgsz=640, batch=32, total=None, metric="images/sec (fwd+NMS, 640^2, bf16)", cpu_batch=32),
    "C3": dict(kind="model", model=SOD, imgsz=1024, batch=16, total=None, metric="images/sec (fwd+NMS, 1024^2, bf16)", cpu_batch=4),
    "C4": dict(kind="model", model=SOD, imgsz=640, batch=MICRO, total=256,
               metric="images/sec (fwd+NMS, 640^2, bf16), 256 images batch-sharded", cpu_batch=8),
    "C4b": dict(kind="model", model="yolov12m", imgsz=640, batch=MICRO, total=256,
                metric="images/sec (fwd+NMS, 640^2, bf16), yolov12m, 256 images batch-sharded", cpu_batch=4),
    "C5u": dict(kind="nms", dist="uniform", A=30000, nc=10, batch=64, total=64,
                metric="images/sec (NMS only, 30k boxes x 10 classes, IoU 0.7, max_det 300)", cpu_batch=1),
    "C5c": dict(kind="nms", dist="clustered", A=30000, nc=10, batch=64, total=64,
                metric="images/sec (NMS only, 30k boxes x 10 classes, clustered, IoU 0.7, max_det 300)", cpu_batch=1),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


# ---- synthetic inputs ---------------------------------------------------------------------------------------------------
def nms_stress_prediction(B, A, nc, dist_name, seed0=0):
    """SURVEY 8d C5 inputs as a (B, 4+nc, A) fp32 prediction tensor: one class id per box, scores a random permutation of A distinct
    values in (0.26, 1) (all pass conf 0.25, no ties); `uniform` boxes barely overlap, `clustered` = jittered clusters (heavy
    suppression). Image b uses seed seed0 + b."""
    import torch
    from tests import nms_cases
    return torch.from_numpy(nms_cases.stress_pred(B, A=A, nc=nc, seed0=seed0, clustered=(dist_name == "clustered")))


# ---- CPU reference legs (oracle port; the one place bench.py executes oracle/) -------------------------------------------
def cpu_reference_leg(cfg_key, steps, warmup, threads=None, budget_s=None):
    """The reference's CPU path (oracle port: fp32 torch CPU ops + the C restatement of torchvision's NMS) on a bounded sample of
    the workload: `cpu_batch` images per step."""
    import torch
    import yolo_sod_b200  # noqa: F401
    from yolo_sod_b200 import cfg as ycfg, synth
    from oracle import model_ref, nms_ref
    c = CONFIGS[cfg_key]
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    nb = c["cpu_batch"]
    if c["kind"] == "model":
        spec = ycfg.get_spec(c["model"])
        sd = synth.synth_state_dict(spec, c["model"], 0)
        strides = ycfg.strides_of(spec)
        x = synth.synth_images(nb, c["imgsz"], seed=0)

        def step():
            y, _ = model_ref.forward(spec, sd, x, strides)
            return nms_ref.non_max_suppression(y.numpy(), CONF, IOU, max_det=MAX_DET)
        what = f"{c['model']} {c['imgsz']}^2 fwd + NMS, fp32 torch CPU ops via oracle/model_ref.py + oracle/nms_ref.py"
    else:
        pred = nms_stress_prediction(nb, c["A"], c["nc"], c["dist"]).numpy()

        def step():
            return nms_ref.non_max_suppression(pred, CONF, IOU, max_det=MAX_DET)
        what = f"NMS only, {c['A']} boxes x {c['nc']} classes ({c['dist']}), oracle/nms_ref.py (C restatement of torchvision's CPU kernel)"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step()
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": nb * done / dt, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"{done} steps x {nb} images of the same workload ({what}), {dt:.1f} s", "ms_per_step": 1e3 * dt / done}


def workload_config(cfg_key, B, world):
    """The `config` object shared by both arms (the reference arm must report the B200 arm's config)."""
    c = CONFIGS[cfg_key]
    if c["kind"] == "model":
        d = {"workload": f"{cfg_key}: {c['model']} fwd+decode+NMS", "imgsz": c["imgsz"]}
    else:
        d = {"workload": f"{cfg_key}: NMS only, {c['A']} candidate boxes x {c['nc']} classes per image ({c['dist']})"}
    total = c["total"] if c["total"] else B * world
    d.update({"batch_per_gpu": B if not c["total"] else -(-c["total"] // world), "global_batch": total, "conf": CONF, "iou": IOU,
              "max_det": MAX_DET, "weights": "synthetic calibrated-random, seed 0",
              "parallelism": f"dp{world} (batch-sharded replicas, one NCCL all_gather of detections per forward)" if world > 1 else "single GPU"})
    return d


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = CONFIGS[args.config]
    leg = cpu_reference_leg(args.config, max(1, args.steps), max(0, args.warmup))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    line = {"impl": "reference", "metric": c["metric"], "value": leg["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if c["total"] else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.config, args.batch or c["batch"], world),
                           note="reference CPU path (oracle port of the pure-Python reference: the same ATen CPU kernels), each step "
                                "= a bounded sample of this workload on all host cores"),
            "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": leg["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---- the reference's stock GPU path on the same B200 (library kernels) ----------------------------------------------------
def gpu_library_leg(cfg_key, dev, steps=5, warmup=2):
    """What the reference itself would run on this GPU: its forward as eager PyTorch (cuDNN / cuBLAS / ATen kernels) in bf16
    channels_last with BN folded (BaseModel.fuse; autobackend.py:154 casts the whole model), then its NMS pipeline with
    torchvision.ops.nms CUDA (ops.py:167-316), same weights and inputs. Runs the oracle port on `cuda`: a baseline leg, never the
    product path."""
    import torch
    import torchvision
    from yolo_sod_b200 import cfg as ycfg, synth
    from oracle import model_ref
    c = CONFIGS[cfg_key]
    B = c["batch"]
    dt = torch.bfloat16

    def ref_nms(y):
        """ops.py:230-297 for the hot-path arguments (multi_label False, no classes filter), batched offsets, torchvision CUDA nms."""
        y = y.float()
        out = []
        xc = y[:, 4:].amax(1) > CONF
        p = y.transpose(-1, -2)
        xy, wh = p[..., :2], p[..., 2:4] / 2
        boxes_all = torch.cat((xy - wh, xy + wh), -1)
        for i in range(y.shape[0]):
            m = xc[i]
            box, cls = boxes_all[i][m], p[i][m][:, 4:]
            conf, j = cls.max(1, keepdim=True)
            x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > CONF]
            if x.shape[0] > 30000:
                x = x[x[:, 4].argsort(descending=True)[:30000]]
            k = torchvision.ops.nms(x[:, :4] + x[:, 5:6] * 7680, x[:, 4], IOU)[:MAX_DET]
            out.append(x[k])
        return out

    if c["kind"] == "model":
        spec = ycfg.get_spec(c["model"])
        sd = model_ref.fuse_state_dict(synth.synth_state_dict(spec, c["model"], 0))
        sd = {k: (v.to(dev, dt) if v.is_floating_point() else v.to(dev)) for k, v in sd.items()}
        for k in list(sd):   # conv weights in channels_last, as `model.to(memory_format=channels_last)` would hold them
            if sd[k].dim() == 4:
                sd[k] = sd[k].contiguous(memory_format=torch.channels_last)
        strides = ycfg.strides_of(spec)
        xs = [synth.synth_images(B, c["imgsz"], seed=i).to(dev).contiguous(memory_format=torch.channels_last) for i in range(2)]

        def step(i):
            y, _ = model_ref.forward(spec, sd, xs[i % 2], strides, dtype=dt)
            return ref_nms(y)
    else:
        preds = [nms_stress_prediction(B, c["A"], c["nc"], c["dist"], seed0=100 * i).to(dev) for i in range(2)]

        def step(i):
            return ref_nms(preds[i % 2])
    with torch.no_grad():
        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            step(i)
        b.record()
        b.synchronize()
    ms = a.elapsed_time(b) / steps
    torch.cuda.empty_cache()
    return {"value": B / ms * 1e3, "unit": "images/s", "ms_per_step": ms, "steps": steps, "dtype": "bf16",
            "what": "reference forward as eager PyTorch on this GPU (cuDNN / cuBLAS / ATen, bf16 channels_last, BN folded) + "
                    "torchvision.ops.nms CUDA per image, via the oracle port on cuda; device-resident inputs, CUDA events"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per forward (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="A/B: run every batch's NMS on the forward's stream (no cross-batch overlap)")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel event timing table to this JSON file")
    ap.add_argument("--quick", action="store_true", help="main timed loop only (for ncu launch lists): no e2e / per-kernel / latency legs")
    args = ap.parse_args()
    _capture_stdout()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import yolo_sod_b200  # noqa: F401
    from yolo_sod_b200 import dist as ydist, ops, synth
    from yolo_sod_b200.model import YOLO

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    c = CONFIGS[args.config]
    is_model = c["kind"] == "model"
    B = args.batch or c["batch"]
    if c["total"]:
        a0, a1 = ydist.shard(c["total"], rank, world)
        mine = a1 - a0                       # images of this rank per step (strong scaling)
        n_micro = -(-mine // B) if mine else 0
        imgs_per_step = c["total"]
        n_micro_max = -(-(-(-c["total"] // world)) // B)
    else:
        mine, n_micro, imgs_per_step, n_micro_max = B, 1, B * world, 1
    IMGSZ = c.get("imgsz", 0)

    side = torch.cuda.Stream(device=dev) if world > 1 else None
    gather = ydist.DetectionGather(world, B, MAX_DET, dev, stream=side) if world > 1 else None
    n_in = 4
    if is_model:
        # overlap_nms: batch i's NMS runs on a side stream under batch i+1's forward (two program slots; model.py)
        yolo = YOLO(c["model"], dtype=torch.bfloat16, device=dev, seed=0, overlap_nms=not args.no_overlap)
        model = yolo.model
        prog = model.program(B, IMGSZ, IMGSZ, False, False, 0)
        nms_stream = yolo.nms_stream
        # rotating device-resident fp32 NCHW input batches (the forward reads them in place; a step touches GBs of activations >> 126 MB L2)
        xs = [synth.synth_images(B, IMGSZ, seed=100 * rank + i).to(dev) for i in range(n_in)]

        def forward_nms(i):
            return yolo.predict_padded(xs[i % n_in], CONF, IOU, MAX_DET)[:2]
        own_launches = prog.n_launches + 1 + 4 + 1   # forward program + input bind + 4 NMS kernels + clip_boxes
    else:
        preds = [nms_stress_prediction(B, c["A"], c["nc"], c["dist"], seed0=1000 * rank + 100 * i).to(dev) for i in range(n_in)]
        prog = None
        nms_stream = None

        def forward_nms(i):
            return ops.nms_padded(preds[i % n_in], CONF, IOU, max_det=MAX_DET)[:2]
        own_launches = 4

    def step(i):
        det = count = None
        for mb in range(n_micro_max):     # every rank issues the same number of collectives
            if mb < n_micro:
                det, count = forward_nms(i * n_micro_max + mb)
            if world > 1:                 # the only data-path exchange: fixed-size detections over NVLink
                if nms_stream is not None:
                    with torch.cuda.stream(nms_stream):   # det / count are produced on the NMS stream: the gather follows them there
                        gather(det, count)
                else:
                    gather(det, count)
        return det, count

    def barrier():
        if is_model:
            yolo.join()
        if gather is not None:
            gather.wait()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        det, count = step(i)
    if is_model:
        yolo.join()                      # the last step's NMS (side stream) is inside the timed region
    if gather is not None:
        gather.wait()                    # ... and so is the last step's gather
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    ms = ydist.max_over_ranks(ms, dev)
    ndet = int(count.sum().item()) if count is not None else 0
    metric = c["metric"]
    if args.quick:
        if rank == 0:
            clocks = sampler.stop() if sampler else None
            emit({"metric": metric, "value": imgs_per_step * args.steps / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
                  "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "quick": True, "clocks": clocks,
                  "config": workload_config(args.config, B, world)})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end-to-end through the public API with HOST buffers ---------------------------------------------------------
    # Model configs: per step the caller hands `YOLO.predict` B raw uint8 BGR frames (B,H,W,3) -- what the reference's predictor
    # hands to preprocess() (engine/predictor.py:116-134) -- copied from pinned host memory on a copy stream (the H2D of step i+1
    # is in flight while predict(i) runs, the usual double buffer), gets back the list of per-image Results, and reads the
    # detections back to host memory. predict() itself ends with a host synchronisation (it returns variable-length tensors).
    # N > 1: the detection all_gather of the step is included. NMS-only configs: H2D of the prediction tensor,
    # ops.non_max_suppression (reference signature), D2H of the detections.
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    gen = torch.Generator().manual_seed(1234 + rank)
    if is_model:
        hx = [torch.randint(0, 256, (B, IMGSZ, IMGSZ, 3), generator=gen, dtype=torch.uint8).pin_memory() for _ in range(2)]
        stage = [torch.empty((B, IMGSZ, IMGSZ, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    else:
        hx = [nms_stress_prediction(B, c["A"], c["nc"], c["dist"], seed0=5000 + 100 * i).pin_memory() for i in range(2)]
        stage = [torch.empty_like(h, device=dev) for h in hx]
    h2d_bytes = hx[0].numel() * hx[0].element_size()
    hdet = [torch.empty((B, MAX_DET, 6), dtype=torch.float32).pin_memory() for _ in range(2)]
    consumed = [0.0]
    pending = []
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_d2h = [torch.cuda.Event() for _ in range(2)]

    def h2d(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[s])          # the forward that read stage[s] two steps ago has consumed it
            stage[s].copy_(hx[s], non_blocking=True)
            ev_in[s].record(copy_stream)

    def e2e_step(i):
        s = i % 2
        h2d(i + 1)                                      # next step's frames travel while this step computes
        main_stream.wait_event(ev_in[s])
        if is_model:
            # the public call; stream=True (engine/model.py:501-560) returns a generator of Results: the batch's device work is
            # enqueued, the host sync happens when the generator is consumed -- one step later, below
            results = yolo.predict(stage[s], stream=True, conf=CONF, iou=IOU, max_det=MAX_DET)
            det_b, cnt_b = results.det, results.count   # the padded batch the per-image Results are views of
            pending.append(results)
            n = 0
            if len(pending) > 1:
                n = sum(len(r) for r in pending.pop(0))   # consume step i-1's Results (host sync on ITS counts) while step i runs
        else:
            rows = ops.non_max_suppression(stage[s], CONF, IOU, max_det=MAX_DET)       # reference signature (host sync inside)
            det_b, cnt_b = rows.det, rows.count
            n = sum(int(r.shape[0]) for r in rows)
        ev_free[s].record(main_stream)
        out_stream = nms_stream if (is_model and nms_stream is not None) else main_stream   # the stream det / count are produced on
        with torch.cuda.stream(out_stream):
            if world > 1:
                gather(det_b, cnt_b)
            # ONE D2H of this step's detections (padded rows; the counts travel separately). Consumed one step later.
            hdet[s].copy_(det_b, non_blocking=True)
            ev_d2h[s].record(out_stream)
        if world > 1:
            gather.wait()
        if i > 0:
            ev_d2h[1 - s].synchronize()
            consumed[0] += float(hdet[1 - s][0, 0, 4])   # the host touches step i-1's detections
        return n

    for s_ in range(2):
        ev_free[s_].record(main_stream)
    h2d(0)
    for i in range(3):
        e2e_step(i)
    barrier()
    k2 = max(4, args.steps)
    t0 = time.perf_counter()
    for i in range(3, 3 + k2):
        e2e_ndet = e2e_step(i)
    while pending:
        e2e_ndet = sum(len(r) for r in pending.pop(0))   # the last step's Results
    main_stream.synchronize()                           # the last step's D2H
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_imgs = B * world * k2    # the e2e leg runs one forward of B images per rank per step (weak), whatever the config's sharding
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    line = {
        "metric": metric, "value": imgs_per_step * args.steps / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if c["total"] else "weak",
        "vs_baseline": None, "dtype": "bf16" if is_model else "f32", "data": "synthetic",
        "config": dict(workload_config(args.config, B, world),
                       l2=f"{n_in} rotating device-resident input batches; a step touches GBs of activations (>> 126 MB L2)" if is_model
                          else f"{n_in} rotating device-resident prediction tensors of {B * 14 * 30000 * 4 / 1e6:.0f} MB + per-image sort workspaces",
                       detections_last_step=ndet),
        "clocks": clocks,
        "e2e": {"value": e2e_imgs / e2e_s, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": B * MAX_DET * 6 * 4 + B * 4,
                "steps": k2, "detections_last_step": e2e_ndet, "batch_per_gpu": B,
                "note": ("YOLO.predict(uint8 BGR HWC frames, stream=True) -- the reference's public call (engine/model.py:501-560) in its "
                         "generator form; frames come from pinned host memory (H2D on a copy stream every step, overlapping the previous "
                         "step), preprocess is fused into the stem, forward + decode + NMS + clip_boxes; the per-image Results of step i-1 "
                         "are consumed (host sync on their counts) and its detections read back (one D2H into pinned host memory) while "
                         "step i runs" if is_model else
                         "ops.non_max_suppression(prediction) with the reference signature: H2D of the (B,14,30000) prediction from pinned "
                         "host memory every step, NMS, host sync (variable-length list), one D2H of the detections")
                        + ("; the step's NCCL all_gather of detections is included" if world > 1 else "")},
        "gpu_launches": args.steps * own_launches * max(n_micro, 1),
        "launches_per_step": own_launches * max(n_micro, 1),
    }

    # ---- per-kernel event timing (eager replay, same stream) -> roofline of the dominant kernel --------------------
    if is_model:
        table = prog.profile(iters=3)
        tot_ms = sum(v["ms"] for v in table.values())
        tc = table.get("ysod_conv_tc_run", {"ms": 0.0, "launches": 0, "flops": 0.0})
        peak_tf = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1400.0)))
        ach_tf = (tc["flops"] / (tc["ms"] * 1e-3)) / 1e12 if tc["ms"] > 0 else 0.0
        traffic = None
        for tname in ("r02_traffic.json", "r01_traffic.json"):   # dram bytes per conv_tc launch, from the committed ncu launch list (C2)
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath) and args.config == "C2":
                traffic = json.load(open(tpath)).get("conv_tc_kernel_dram_bytes_per_launch")
                break
        line["roofline"] = {
            "bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv/linear)", "achieved": round(ach_tf, 2),
            "peak": peak_tf, "unit": "TFLOP/s", "frac": round(ach_tf / peak_tf, 4), "traffic": traffic,
            "traffic_unit": "bytes of DRAM read+write per launch (ncu, average over the launches of one step)",
            "algorithmic_gflop_per_launch": round(tc["flops"] / 1e9 / max(tc["launches"], 1), 2),
            "peak_source": f"{pk_src} bf16_tflops_sustained (kernel timed inside a long step)",
            "launches_per_step": tc["launches"], "algorithmic_gflop_per_step": round(tc["flops"] / 1e9, 2),
            "share_of_step": round(tc["ms"] / tot_ms, 4) if tot_ms else None, "ms_per_step_in_kernel": round(tc["ms"], 3)}
        if args.profile_out:
            os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
            json.dump({"config": args.config, "batch": B, "imgsz": IMGSZ, "sum_ms": tot_ms, "kernels": table, "per_op": prog.last_per_op},
                      open(args.profile_out, "w"), indent=1)

        def timed(fn, n=10):
            fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(n):
                fn()
            b.record()
            b.synchronize()
            return a.elapsed_time(b) / n
        y_last, _ = model(xs[0], static=True, want_raw=False)
        line["step_split_ms"] = {"forward_graph_ms": round(timed(lambda: model(xs[1], static=True, want_raw=False)), 4),
                                 "nms_ms": round(timed(lambda: ops.nms_padded(y_last, CONF, IOU, max_det=MAX_DET)), 4),
                                 "forward_with_raw_maps_ms": round(timed(lambda: model(xs[1], static=True)), 4)}
        # ---- batch-1 latency (second half of the BASELINE metric) -----------------------------------------------------
        x1 = synth.synth_images(1, IMGSZ, seed=7).to(dev)
        for _ in range(5):
            yolo.predict_padded(x1, CONF, IOU, MAX_DET)
        torch.cuda.synchronize()
        lat = []
        for _ in range(100):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            yolo.predict_padded(x1, CONF, IOU, MAX_DET)
            yolo.join()                      # the batch's NMS (side stream in overlap mode) belongs to its latency
            b.record()
            b.synchronize()
            lat.append(a.elapsed_time(b))
        lat.sort()
        line["latency_b1_ms"] = {"p50": lat[len(lat) // 2], "p90": lat[int(len(lat) * 0.9)], "min": lat[0], "runs": len(lat)}
    else:
        # NMS: HBM-side algorithmic bytes per image = prediction read (4+nc)*A*4 + sorted boxes/scores/classes written and read
        # once (24 B per candidate) + output rows; the kept-list sweep replaces SURVEY 8d's n^2/16 B IoU mask (never materialised).
        alg = (4 + c["nc"]) * c["A"] * 4 + 2 * 24 * c["A"] + MAX_DET * 28
        hbm = float(pk.get("hbm_gbs", 6545.9))
        ach = alg * imgs_per_step / world * args.steps / (ms * 1e-3) / 1e9
        line["roofline"] = {"bound": "hbm", "kernel": "nms_score / nms_sort / nms_gather / nms_chunk (whole NMS pipeline)",
                            "achieved": round(ach, 1), "peak": hbm, "unit": "GB/s", "frac": round(ach / hbm, 4), "traffic": None,
                            "algorithmic_bytes_per_image": alg, "peak_source": f"{pk_src} hbm_gbs",
                            "note": "latency-bound by the sequential greedy sweep (dependent chunks of 32 candidates), not by bytes"}
    if world == 1 and not args.no_library_baseline:
        try:
            line["gpu_library_baseline"] = gpu_library_leg(args.config, dev)
        except Exception as e:   # a baseline leg must never take the measurement down
            line["gpu_library_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    if world == 1 and not args.no_cpu_baseline:
        ncores = os.cpu_count() or 1
        cpu = cpu_reference_leg(args.config, 64, 1, threads=ncores, budget_s=10.0)
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if ncores != 8:   # the reference's own convention: NUM_THREADS = min(8, cpu_count - 1) (utils/__init__.py:43)
            cpu8 = cpu_reference_leg(args.config, 64, 1, threads=min(8, ncores), budget_s=10.0)
            line["cpu_baseline_t8"] = {k: cpu8[k] for k in ("value", "unit", "cores", "kind", "sample")}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
