#!/usr/bin/env python
"""Benchmark of the hot path: YOLOv12-SOD forward + decode + NMS, 640x640, bf16, images/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path (forward graph + decode + batched NMS) over one batch of synthetic images.
N=1 workload = BASELINE.json configs[1]: SOD fusion v5-simple, 640^2, batch 32 per GPU, bf16. N>1: one process per GPU
(torchrun), the image batch is sharded (weak scaling: 32 images per GPU), the only data-path exchange is an NCCL all_gather
of the padded detections. Rank 0 prints one JSON line.

`--impl reference` times the reference's own CPU implementation of the same path on the host cores. The reference is pure
Python on torch CPU ops and cannot be installed on the GPU box (its package is incomplete), so the arm
runs the oracle port (oracle/model_ref.py + oracle/nms_ref.py: the same ATen CPU kernels, pinned bit-close to the live
reference) on a bounded sample of the same workload.
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

CFG = "yolov12-sod-fusion-v5-simple"
IMGSZ = 640
BATCH = 32
CONF, IOU, MAX_DET = 0.25, 0.7, 300
METRIC = "images/sec (fwd+NMS, 640^2, bf16)"


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


def cpu_reference_leg(steps, warmup, sample_batch=BATCH, threads=None):
    """The reference's CPU path (oracle port) on a bounded sample: forward + NMS on `sample_batch` 640^2 images per step
    (default: one full batch of the workload, ~1.5 s on 24 cores)."""
    import torch
    import yolo_sod_b200  # noqa: F401
    from yolo_sod_b200 import cfg as ycfg, synth
    from oracle import model_ref, nms_ref
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    spec = ycfg.get_spec(CFG)
    sd = synth.synth_state_dict(spec, CFG, 0)
    strides = ycfg.strides_of(spec)
    x = synth.synth_images(sample_batch, IMGSZ, seed=0)

    def step():
        y, _ = model_ref.forward(spec, sd, x, strides)
        return nms_ref.non_max_suppression(y.numpy(), CONF, IOU, max_det=MAX_DET)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": sample_batch * steps / dt, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"{steps} steps x {sample_batch} images of the same workload (SOD 640^2 fwd + NMS, fp32 torch CPU "
                      f"ops via oracle/model_ref.py + oracle/nms_ref.py), {dt:.1f} s",
            "ms_per_step": 1e3 * dt / steps}


def workload_config(B, world):
    """The `config` object shared by both arms (the reference arm must report the B200 arm's config)."""
    return {"workload": f"{CFG} fwd+decode+NMS", "imgsz": IMGSZ, "batch_per_gpu": B, "global_batch": B * world,
            "conf": CONF, "iou": IOU, "max_det": MAX_DET, "weights": "synthetic calibrated-random, seed 0",
            "parallelism": f"dp{world} (batch-sharded replicas, NCCL all_gather of detections)" if world > 1 else "single GPU"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    leg = cpu_reference_leg(max(1, args.steps), max(0, args.warmup))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    line = {"impl": "reference", "metric": METRIC, "value": leg["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.batch, world), note="reference CPU path (oracle port of the pure-Python reference: the "
                           "same ATen CPU kernels), each step = one batch of this workload on all host cores"),
            "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": leg["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="images per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from yolo_sod_b200.model import DetectionModel

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

    B = args.batch
    model = DetectionModel(CFG, dtype=torch.bfloat16, device=dev, seed=0)
    prog = model.program(B, IMGSZ, IMGSZ)
    # 4 rotating device-resident input batches (each 157 MB fp32; a step touches ~5 GB of activations >> 126 MB L2)
    n_in = 4
    xs = [synth.synth_images(B, IMGSZ, seed=100 * rank + i).to(dev) for i in range(n_in)]
    gather = ydist.DetectionGather(world, B, MAX_DET, dev) if world > 1 else None   # preallocated (world*B, 300, 6) + (world*B,) buffers

    def step(i):
        y, _ = model(xs[i % n_in])
        det, count, _ = ops.nms_padded(y, CONF, IOU, max_det=MAX_DET)
        if world > 1:  # the only data-path exchange: fixed-size detections over NVLink
            gather(det, count)
        return det, count

    def barrier():
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
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    ms = ydist.max_over_ranks(ms, dev)
    ndet = int(count.sum().item())
    if args.quick:
        if rank == 0:
            clocks = sampler.stop() if sampler else None
            emit({"metric": METRIC, "value": world * B * args.steps / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
                  "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "quick": True, "clocks": clocks})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end-to-end through the public API with HOST buffers ---------------------------------------------------------
    # Input per step: B raw uint8 BGR frames (B,640,640,3) in pinned host memory -- what the reference's predictor hands to
    # preprocess() (engine/predictor.py:116-134); BGR->RGB / HWC->CHW / /255 are fused into the stem kernel. Every step does
    # its own H2D copy (copy stream, double-buffered so it overlaps the previous step's kernels) and its own D2H read of the
    # detections into pinned host memory; the host waits for step i-1's result while step i runs (lag-1 pipeline).
    from yolo_sod_b200.model import YOLO  # noqa: F401  (public surface; predict() wraps the same two calls)
    gen = torch.Generator().manual_seed(1234 + rank)
    hx = [torch.randint(0, 256, (B, IMGSZ, IMGSZ, 3), generator=gen, dtype=torch.uint8).pin_memory() for _ in range(2)]
    stage = [torch.empty((B, IMGSZ, IMGSZ, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    hdet = [torch.empty((B, MAX_DET, 6), dtype=torch.float32).pin_memory() for _ in range(2)]
    hcnt = [torch.empty((B,), dtype=torch.int32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def e2e_submit(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[s])                  # the forward that read stage[s] two steps ago has consumed it
            stage[s].copy_(hx[s], non_blocking=True)            # H2D of this step's frames
            ev_in[s].record(copy_stream)
        main_stream.wait_event(ev_in[s])
        y, _ = model(stage[s])                                  # DetectionModel.forward (uint8 frames) -> (y, raw)
        ev_free[s].record(main_stream)
        d, c, _ = ops.nms_padded(y, CONF, IOU, max_det=MAX_DET)
        hdet[s].copy_(d, non_blocking=True)                     # D2H of this step's detections
        hcnt[s].copy_(c, non_blocking=True)
        ev_out[s].record(main_stream)

    for s_ in range(2):
        ev_free[s_].record(main_stream)
    for i in range(3):
        e2e_submit(i)
    barrier()
    k2 = max(4, args.steps)
    t0 = time.perf_counter()
    for i in range(k2):
        e2e_submit(i)
        if i > 0:
            ev_out[(i - 1) % 2].synchronize()                  # the caller consumes step i-1's detections
    ev_out[(k2 - 1) % 2].synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_ndet = int(hcnt[(k2 - 1) % 2].sum())
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel event timing (eager replay, same stream) -> roofline of the dominant kernel --------------------
    table = prog.profile(iters=3)
    tot_ms = sum(v["ms"] for v in table.values())
    tc = table.get("ysod_conv_tc_run", {"ms": 0.0, "launches": 0, "flops": 0.0})
    pk, pk_src = peaks()
    peak_tf = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1400.0)))
    ach_tf = (tc["flops"] / (tc["ms"] * 1e-3)) / 1e12 if tc["ms"] > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")   # dram__bytes_read+write per conv_tc launch, from the committed ncu launch list
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("conv_tc_kernel_dram_bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv/linear)", "achieved": round(ach_tf, 2),
                "peak": peak_tf, "unit": "TFLOP/s", "frac": round(ach_tf / peak_tf, 4), "traffic": traffic,
                "traffic_unit": "bytes of DRAM read+write per launch (ncu, average over the launches of one step)",
                "algorithmic_gflop_per_launch": round(tc["flops"] / 1e9 / max(tc["launches"], 1), 2),
                "peak_source": f"{pk_src} bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_per_step": tc["launches"], "algorithmic_gflop_per_step": round(tc["flops"] / 1e9, 2),
                "share_of_step": round(tc["ms"] / tot_ms, 4) if tot_ms else None,
                "ms_per_step_in_kernel": round(tc["ms"], 3)}
    if args.profile_out:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
        json.dump({"batch": B, "imgsz": IMGSZ, "sum_ms": tot_ms, "kernels": table, "per_op": prog.last_per_op},
                  open(args.profile_out, "w"), indent=1)

    # ---- split of the step: forward graph alone vs NMS alone (device events, same stream, 10 runs each)
    def timed(fn, n=10):
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / n
    y_last, _ = model(xs[0])
    split = {"forward_graph_ms": round(timed(lambda: model(xs[1])), 4),
             "nms_ms": round(timed(lambda: ops.nms_padded(y_last, CONF, IOU, max_det=MAX_DET)), 4)}

    # ---- batch-1 latency (second half of the BASELINE metric) -----------------------------------------------------
    x1 = synth.synth_images(1, IMGSZ, seed=7).to(dev)
    for _ in range(5):
        ops.nms_padded(model(x1)[0], CONF, IOU, max_det=MAX_DET)
    torch.cuda.synchronize()
    lat = []
    for _ in range(100):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.nms_padded(model(x1)[0], CONF, IOU, max_det=MAX_DET)
        b.record()
        b.synchronize()
        lat.append(a.elapsed_time(b))
    lat.sort()

    cpu = None if args.no_cpu_baseline else cpu_reference_leg(8, 1)   # ~10-15 s of CPU work
    n_nms = 4
    imgs = world * B * args.steps
    line = {
        "metric": METRIC, "value": imgs / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": dict(workload_config(B, world),
                       l2="4 rotating input batches of 157 MB; ~2 GB of activations touched per step (>> 126 MB L2)",
                       detections_last_step=ndet),
        "clocks": clocks,
        "e2e": {"value": world * B * k2 / e2e_s, "unit": "images/s", "h2d_bytes_per_step": B * 3 * IMGSZ * IMGSZ,
                "d2h_bytes_per_step": B * MAX_DET * 6 * 4 + B * 4, "steps": k2, "detections_last_step": e2e_ndet,
                "note": "uint8 BGR HWC frames (the predictor's raw input, predictor.py:116-134) from pinned host memory, H2D on a "
                        "copy stream every step, DetectionModel.forward (preprocess fused into the stem) + ops.nms_padded, "
                        "detections D2H to pinned host memory every step; host consumes step i-1 while step i runs"},
        "gpu_launches": args.steps * (prog.n_launches + n_nms),
        "launches_per_step": prog.n_launches + n_nms,
        "roofline": roofline,
        "step_split_ms": split,
        "latency_b1_ms": {"p50": lat[len(lat) // 2], "p90": lat[int(len(lat) * 0.9)], "min": lat[0], "runs": len(lat)},
    }
    if cpu:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
