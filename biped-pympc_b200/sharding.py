"""Env-dimension sharding across the GPUs of one box.

Every environment is an independent QP, so the batch is split into contiguous slices, one
per rank (one process per GPU, torch.distributed for the plumbing) and there is **no
collective on the solve path**.  The only optional exchange is a gather of the
``(B/G, 2, 6)`` fp32 wrenches for a single-process consumer."""
from __future__ import annotations

from typing import List, Tuple


def shard_bounds(num_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the contiguous env slice owned by ``rank`` (remainder spread over the
    first ranks, so slice sizes differ by at most one)."""
    if world_size <= 0 or not (0 <= rank < world_size) or num_envs < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(num_envs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedSolver:
    """Owns this rank's slice of a ``num_envs`` batch and (optionally) gathers results.

    ``solve_fn(inputs17_slice, rotation_slice) -> wrench_slice`` is the per-rank solve
    (on a GPU: ``Context.solve``).  ``gather`` uses ``torch.distributed.all_gather`` (NCCL on
    GPUs, gloo in the CPU tests); it is never called on the timed solve path."""

    def __init__(self, num_envs: int, solve_fn, world_size: int = None, rank: int = None):
        import torch.distributed as dist
        self._dist = dist if dist.is_available() and dist.is_initialized() else None
        self.world_size = world_size if world_size is not None else (self._dist.get_world_size() if self._dist else 1)
        self.rank = rank if rank is not None else (self._dist.get_rank() if self._dist else 0)
        self.num_envs = num_envs
        self.lo, self.hi = shard_bounds(num_envs, self.world_size, self.rank)
        self.solve_fn = solve_fn

    def local(self, full_tensors: List):
        return [t[self.lo:self.hi] for t in full_tensors]

    def solve(self, inputs17_local, rotation_local):
        return self.solve_fn(inputs17_local, rotation_local)

    def gather(self, wrench_local):
        """(num_envs, 2, 6) on every rank.  Slices may differ by one env, so pad to the max."""
        import torch
        if self._dist is None or self.world_size == 1:
            return wrench_local
        sizes = [shard_bounds(self.num_envs, self.world_size, r) for r in range(self.world_size)]
        mx = max(hi - lo for lo, hi in sizes)
        pad = torch.zeros((mx,) + tuple(wrench_local.shape[1:]), dtype=wrench_local.dtype, device=wrench_local.device)
        pad[: wrench_local.shape[0]] = wrench_local
        parts = [torch.empty_like(pad) for _ in range(self.world_size)]
        self._dist.all_gather(parts, pad)
        return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
