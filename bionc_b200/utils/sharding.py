"""Multi-GPU plumbing: the batch of independent systems is partitioned contiguously over the ranks
(one process per GPU); there is NO collective on the step path.  The only communication is the
device-time reduction used for reporting and an optional final gather of the states
(``torch.distributed``: NCCL over NVLink on GPUs, gloo in the CPU tests)."""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition of ``range(total)``: rank ``r`` owns ``[lo, hi)``; the first
    ``total % world`` ranks hold one extra system."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(values, device=None) -> list[float]:
    """Element-wise maximum of a list of floats over all ranks (slowest rank defines the job time)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def gather_states(local: torch.Tensor, total: int) -> torch.Tensor:
    """Final gather: every rank contributes its ``(b_r, d)`` shard (shards may differ by one row) and
    receives the full ``(total, d)`` tensor in system order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * width, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * width : r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)
