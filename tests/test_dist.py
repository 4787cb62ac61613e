"""N > 1 path on CPU (gloo, world size 2): batch sharding and the detection all_gather of yolo_sod_b200/dist.py -- the only
data-path exchange of the multi-GPU run (bench.py --gpus N uses the same class over NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import yolo_sod_b200  # noqa: F401
from yolo_sod_b200 import dist as ydist

MAX_DET = 300


def _fake_detections(n_images, seed=0):
    g = torch.Generator().manual_seed(seed)
    det = torch.rand((n_images, MAX_DET, 6), generator=g)
    count = torch.randint(0, MAX_DET + 1, (n_images,), generator=g, dtype=torch.int32)
    return det, count


def _worker(rank, world, port, n_images, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        det_all, count_all = _fake_detections(n_images)
        per = -(-n_images // world)                              # every rank contributes `per` rows (last shard padded)
        a, b = ydist.shard(n_images, rank, world)
        det = torch.zeros((per, MAX_DET, 6))
        count = torch.zeros((per,), dtype=torch.int32)
        det[: b - a] = det_all[a:b]
        count[: b - a] = count_all[a:b]
        gather = ydist.DetectionGather(world, per, MAX_DET, "cpu")
        calls = []
        real = dist.all_gather_into_tensor
        dist.all_gather_into_tensor = lambda *a, **k: (calls.append(1), real(*a, **k))[1]
        for _ in range(2):                                       # buffers are reused across steps
            gd, gc = gather(det, count)
        dist.all_gather_into_tensor = real
        assert len(calls) == 2, "one collective per step: the count travels inside the detection buffer"
        assert gc.dtype == torch.int32
        ms = ydist.max_over_ranks(10.0 + rank, "cpu")
        torch.save({"det": gd.clone(), "count": gc.clone(), "ms": ms, "range": (a, b)}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_is_a_contiguous_partition():
    for n, w in [(256, 8), (33, 2), (5, 8), (0, 4), (32, 1)]:
        r = [ydist.shard(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@pytest.mark.parametrize("n_images", [8, 7])
def test_detection_gather_world2_gloo(tmp_path, n_images):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n_images, str(tmp_path)), nprocs=2, join=True)
    det_all, count_all = _fake_detections(n_images)
    per = -(-n_images // 2)
    outs = [torch.load(tmp_path / f"r{k}.pt") for k in range(2)]
    assert outs[0]["ms"] == outs[1]["ms"] == 11.0                # max over ranks
    for o in outs:                                               # every rank holds the whole job's detections, in rank order
        for k in range(2):
            a, b = outs[k]["range"]
            assert torch.equal(o["det"][k * per: k * per + (b - a)], det_all[a:b])
            assert torch.equal(o["count"][k * per: k * per + (b - a)], count_all[a:b])
        assert int(o["count"].sum()) == int(count_all.sum())    # padded rows carry count 0
