"""Multi-GPU plumbing: one process per GPU, torch.distributed for the rendezvous and the ONE
collective the path has.

The path shards two ways, both without a data-path collective:
  * portfolio on one terrain — rank r runs chains [r*n, (r+1)*n) (independent counter-based RNG streams); once per
    epoch the ranks share the best-known support count with a single all-reduce-min (4 bytes: latency-bound,
    NVLink bandwidth is irrelevant) so every chain only looks for strictly better layouts;
  * terrain batch — contiguous ranges of terrains per rank, no exchange until the final gather of the counts.
Nothing here computes: the search/evaluation objects are the engine's (libtss); tests drive this module on CPU
with the gloo backend and a scripted stand-in for the device search.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

NO_BOUND = 1 << 20


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of n units for `rank` (first n % world ranks get one extra)."""
    if world <= 0 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def allreduce_min(value: Optional[int], device=None) -> Optional[int]:
    """all-reduce-min of the best-known count over the ranks; None = this rank has no complete layout yet."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([NO_BOUND if value is None else int(value)], dtype=torch.int32, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    v = int(t.item())
    return None if v >= NO_BOUND else v


class Portfolio:
    """Independent-seed SLS portfolio on one terrain across the ranks of the default process group."""

    def __init__(self, make_search, chains_per_rank: int, device=None, native: bool = False):
        """make_search(chain_offset, n_chains) -> an object with run(steps, target), best_count(), set_bound(c),
        best_layout() — Engine.search(...) on a GPU rank.

        native=True: the engine has a communicator (Engine.comm_init) and all-reduce-mins the device-resident bound
        itself inside run(); this class then only reads the global best.  native=False: the exchange goes through
        torch.distributed (any backend — what the gloo tests exercise)."""
        dist = _dist()
        self.rank = dist.get_rank() if dist else 0
        self.world = dist.get_world_size() if dist else 1
        self.device = device
        self.native = native
        self.search = make_search(self.rank * chains_per_rank, chains_per_rank)
        self.global_best: Optional[int] = None
        self.epochs = 0

    def epoch(self, steps: int, target: int = 0) -> Optional[int]:
        """One epoch on every rank followed by the all-reduce-min; returns the global best count."""
        self.search.run(steps, target)
        if self.native:
            best = self.search.global_best()   # the engine already exchanged the bound in-stream (ncclAllReduce min)
        else:
            best = allreduce_min(self.search.best_count(), self.device)
            if best is not None:
                self.search.set_bound(best)    # chains now only look for layouts with fewer than `best` supports
        if best is not None:
            self.global_best = best if self.global_best is None else min(self.global_best, best)
        self.epochs += 1
        return self.global_best

    def owner_rank(self) -> Optional[int]:
        """Lowest rank holding a layout with the global best count (all-reduce-min of rank-or-inf)."""
        local = self.search.best_count()
        mine = self.rank if (local is not None and local == self.global_best) else None
        return allreduce_min(mine, self.device)


def solve_batch_sharded(solve_local, grids: np.ndarray, device=None) -> np.ndarray:
    """Terrain batch: rank r solves its contiguous range with solve_local(grids[lo:hi], first_index=lo) -> counts,
    then the per-terrain counts are gathered on every rank (the only exchange: 4 bytes per terrain)."""
    dist = _dist()
    n = len(grids)
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    lo, hi = shard_range(n, rank, world)
    local = np.asarray(solve_local(grids[lo:hi], lo), dtype=np.int32)
    if world == 1:
        return local
    import torch
    sizes = [shard_range(n, r, world) for r in range(world)]
    width = max(h - l for l, h in sizes)
    buf = torch.full((width,), -1, dtype=torch.int32, device=device or "cpu")
    buf[: hi - lo] = torch.from_numpy(local).to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return np.concatenate([o[: h - l].cpu().numpy() for o, (l, h) in zip(out, sizes)])
