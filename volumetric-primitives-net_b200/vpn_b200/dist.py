"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The hot path shards by batch with no data-path collective: every sample's
sampling, Chamfer, mesh and silhouette is independent and the only cross-sample operation is the final
mean over B (chamfer_distance.py:30; L1Loss mean, silhouette.py:11,22).  The one collective per step is
the all-reduce of the network-parameter gradients, which the reference does not have (it is single
process); it is issued on a side stream so it overlaps whatever the caller runs next.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Rank r of G owns samples [lo, hi); shards differ by at most one sample."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_loss_scale(global_batch: int, rank: int, world: int) -> float:
    """Factor that turns a rank's local batch-mean into its share of the global batch-mean, so that the
    SUM all-reduce of the gradients equals the single-process gradient."""
    lo, hi = shard_range(global_batch, rank, world)
    return (hi - lo) / float(global_batch)


class GradientAllReduce:
    """Sum all-reduce of one flat gradient buffer on a dedicated stream (NCCL), joined on demand."""

    def __init__(self, numel: int, device, dtype=torch.float32):
        self.buf = torch.zeros(numel, dtype=dtype, device=device)
        self.cuda = torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.done: Optional[torch.cuda.Event] = None

    def launch(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        if not self.cuda:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
            return
        self.stream.wait_stream(torch.cuda.current_stream(self.buf.device))
        with torch.cuda.stream(self.stream):
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
            self.done = torch.cuda.Event()
            self.done.record(self.stream)

    def join(self):
        if self.done is not None:
            torch.cuda.current_stream(self.buf.device).wait_event(self.done)
            self.done = None
