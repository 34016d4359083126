"""Batch sharding across GPUs: one process per GPU, contiguous split of dim 0, no
data-path collective.  The only exchange on the path is an all-gather of the per-item loudness vector
(``[B/W] f32`` per rank -> ``[B]``), for whole-batch loudness statistics / logging; with NCCL it is issued
on a side stream so that it overlaps the spectral kernel.  Works with any ``torch.distributed`` backend
(NCCL over NVLink on the B200 box, gloo in the CPU tests)."""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch_size: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Items ``[lo, hi)`` owned by ``rank``: contiguous, sizes differ by at most one, in rank order."""
    base, rem = divmod(batch_size, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x, rank: Optional[int] = None, world_size: Optional[int] = None):
    """Slice a ``[B, ...]`` tensor or an ``AudioSignal`` along the batch dim for this rank."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    n = x.batch_size if hasattr(x, "batch_size") else x.shape[0]
    lo, hi = shard_bounds(n, rank, world_size)
    return x[lo:hi]


class LoudnessGather:
    """All-gather of per-item loudness across ranks, optionally on a CUDA side stream."""

    def __init__(self, group=None, side_stream: Optional["torch.cuda.Stream"] = None):
        self.group = group
        self.side = side_stream
        self._out = None

    def __call__(self, loud_local: torch.Tensor, counts=None) -> torch.Tensor:
        """``loud_local`` [B_local] -> [B_total] (equal shard sizes, or ``counts`` per rank)."""
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return loud_local
        world = dist.get_world_size(self.group)
        if counts is None or len(set(counts)) == 1:
            out = torch.empty(world * loud_local.numel(), dtype=loud_local.dtype, device=loud_local.device)
            if self.side is not None and loud_local.is_cuda:
                self.side.wait_stream(torch.cuda.current_stream(loud_local.device))
                with torch.cuda.stream(self.side):
                    dist.all_gather_into_tensor(out, loud_local.contiguous(), group=self.group)
                loud_local.record_stream(self.side)
            else:
                dist.all_gather_into_tensor(out, loud_local.contiguous(), group=self.group)
            return out
        # uneven shards: pad to the largest, gather, drop the padding (all_gather needs equal sizes)
        cmax = max(counts)
        padded = torch.zeros(cmax, dtype=loud_local.dtype, device=loud_local.device)
        padded[: loud_local.numel()] = loud_local
        out = torch.empty(world * cmax, dtype=loud_local.dtype, device=loud_local.device)
        dist.all_gather_into_tensor(out, padded, group=self.group)
        return torch.cat([out[r * cmax: r * cmax + c] for r, c in enumerate(counts)])

    def wait(self):
        """Make the current stream wait for a side-stream gather before its result is consumed."""
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
