"""Data-parallel plumbing of the hot path: images are independent end to end, so a batch is split
contiguously across ranks (one process per GPU, weights replicated, no collective inside the forward) and the only exchange is
an all_gather of the fixed-size padded detections `(B/G, max_det, 6) fp32 + (B/G,) int32` (7.2 KB per image) -- NCCL over NVLink on
the GPU box, gloo in the CPU tests. Nothing in the reference corresponds to this (it has no multi-GPU inference path)."""
from typing import Tuple

import torch
import torch.distributed as dist


def shard(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `n_items` owned by `rank`: the first `n_items % world` ranks take one extra item."""
    base, extra = divmod(int(n_items), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class DetectionGather:
    """all_gather of the padded NMS output with preallocated receive buffers (no allocation / no host sync per step).
    Every rank must contribute the same per-rank batch `B` (pad the last shard); `counts` tells how many rows of each image are valid."""

    def __init__(self, world: int, B: int, max_det: int, device, group=None):
        self.world, self.B, self.group = world, B, group
        # concatenated layout (world*B, ...): the form of all_gather_into_tensor that both NCCL and gloo accept
        self.det = torch.empty((world * B, max_det, 6), dtype=torch.float32, device=device)
        self.count = torch.empty((world * B,), dtype=torch.int32, device=device)

    def __call__(self, det: torch.Tensor, count: torch.Tensor):
        """det (B, max_det, 6) fp32, count (B,) int32 of this rank -> (world*B, max_det, 6), (world*B,) in rank order."""
        if self.world > 1:
            dist.all_gather_into_tensor(self.det, det.contiguous(), group=self.group)
            dist.all_gather_into_tensor(self.count, count.contiguous(), group=self.group)
        else:
            self.det.copy_(det)
            self.count.copy_(count)
        return self.det, self.count


def max_over_ranks(value: float, device, group=None) -> float:
    """Timing rule of the bench contract: a multi-GPU number is the max over ranks of the device-timed duration."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
