"""Data-parallel plumbing of the hot path: images are independent end to end, so a batch is split
contiguously across ranks (one process per GPU, weights replicated, no collective inside the forward) and the only exchange is
ONE all_gather of the fixed-size padded detections `(B/G, max_det, 6) fp32` with the per-image count packed into the same buffer
(7.2 KB per image) -- NCCL over NVLink on the GPU box, gloo in the CPU tests. Nothing in the reference corresponds to this (it has no multi-GPU inference path)."""
from typing import Tuple

import torch
import torch.distributed as dist


def shard(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `n_items` owned by `rank`: the first `n_items % world` ranks take one extra item."""
    base, extra = divmod(int(n_items), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class DetectionGather:
    """all_gather of the padded NMS output with preallocated buffers (no allocation / no host sync per step).

    One collective per step: each image contributes `max_det + 1` rows of 6 floats -- its `max_det` detection rows and one
    trailer row whose first element carries the int32 count bit-cast to fp32 (exact: it is never converted) -- so the count
    travels with the boxes instead of in a second, equally latency-bound all_gather. With `stream=` the pack + collective are
    issued on a side stream behind an event, so step i's gather overlaps step i+1's forward; `wait()` joins it.
    Every rank must contribute the same per-rank batch `B` (pad the last shard); `counts` tells how many rows of each image are valid."""

    def __init__(self, world: int, B: int, max_det: int, device, group=None, stream=None):
        self.world, self.B, self.max_det, self.group = world, B, max_det, group
        self.stream = stream
        # concatenated layout (world*B, ...): the form of all_gather_into_tensor that both NCCL and gloo accept
        self.send = torch.zeros((B, max_det + 1, 6), dtype=torch.float32, device=device)
        self.recv = torch.empty((world * B, max_det + 1, 6), dtype=torch.float32, device=device)
        self.det = self.recv[:, :max_det]                     # (world*B, max_det, 6) view
        self._done = None

    @property
    def count(self):
        """(world*B,) int32, a fresh tensor decoded from the trailer rows."""
        return self.recv[:, self.max_det, :1].contiguous().view(torch.int32).reshape(-1)

    def _issue(self, det: torch.Tensor, count: torch.Tensor):
        self.send[:, :self.max_det].copy_(det)
        self.send[:, self.max_det, 0].copy_(count.view(torch.float32))    # bit pattern, not a numeric conversion
        if self.world > 1:
            dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        else:
            self.recv.copy_(self.send)

    def __call__(self, det: torch.Tensor, count: torch.Tensor):
        """det (B, max_det, 6) fp32, count (B,) int32 of this rank -> (world*B, max_det, 6), (world*B,) in rank order.
        Side-stream mode: returns immediately with the buffers that `wait()` makes valid; the caller must not reuse `det` /
        `count` storage before then (bench: they are fresh tensors each step)."""
        if self.stream is None or not det.is_cuda:
            self._issue(det, count)
            return self.det, self.count
        cur = torch.cuda.current_stream(det.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self._issue(det, count)
            det.record_stream(self.stream)
            count.record_stream(self.stream)
            self._done = torch.cuda.Event()
            self._done.record(self.stream)
        return self.det, None

    def wait(self):
        """Joins the side-stream gather into the current stream; returns (det, count)."""
        if self._done is not None:
            torch.cuda.current_stream(self.recv.device).wait_event(self._done)
            self._done = None
        return self.det, self.count


def max_over_ranks(value: float, device, group=None) -> float:
    """Timing rule of the bench contract: a multi-GPU number is the max over ranks of the device-timed duration."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
