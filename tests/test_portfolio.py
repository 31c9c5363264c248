"""N > 1 host logic on CPU: world_size-2 gloo process groups drive timberborn_support_solver_b200.portfolio with a
scripted stand-in for the device search (no compute happens here — the GPU path has no CPU fallback)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from timberborn_support_solver_b200 import portfolio as P


def test_shard_range_balanced_and_contiguous():
    for n in (0, 1, 7, 8, 100000):
        for world in (1, 2, 3, 8):
            parts = [P.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [h - l for l, h in parts]
            assert max(sizes) - min(sizes) <= 1
    assert P.shard_range(100000, 3, 8) == (37500, 50000)      # 12 500 terrains per GPU
    with pytest.raises(ValueError):
        P.shard_range(10, 2, 2)


class ScriptedSearch:
    """Stand-in for Engine.search(): best counts follow a script; records the bounds it is given."""

    def __init__(self, chain_offset, n_chains, script):
        self.chain_offset, self.n_chains, self.script, self.epoch, self.bounds = chain_offset, n_chains, script, -1, []

    def run(self, steps, target):
        self.epoch += 1

    def best_count(self):
        return self.script[min(self.epoch, len(self.script) - 1)]

    def set_bound(self, c):
        self.bounds.append(c)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scripts = {0: [None, 18, 17, 17], 1: [20, 19, 16, 16]}
        pf = P.Portfolio(lambda off, n: ScriptedSearch(off, n, scripts[rank]), chains_per_rank=64)
        assert pf.search.chain_offset == rank * 64 and pf.world == world and pf.rank == rank
        bests = [pf.epoch(100) for _ in range(4)]
        owner = pf.owner_rank()
        # terrain batch: rank r "solves" its range by returning first_index + i, gathered everywhere
        grids = np.zeros((11, 2, 2), np.uint8)
        counts = P.solve_batch_sharded(lambda g, lo: np.arange(lo, lo + len(g)), grids)
        out.put((rank, bests, pf.search.bounds, owner, counts.tolist(), P.allreduce_min(None)))
    finally:
        dist.destroy_process_group()


def test_portfolio_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, bests, bounds, owner, counts, none_min in res:
        assert bests == [20, 18, 16, 16]            # min over ranks each epoch, never increasing
        assert bounds == [20, 18, 16, 16]           # every rank adopts the global bound
        assert owner == 1                           # rank 1 holds the 16
        assert counts == list(range(11))            # contiguous shards, gathered in order on every rank
        assert none_min is None


def test_single_process_paths():
    pf = P.Portfolio(lambda off, n: ScriptedSearch(off, n, [None, 5]), chains_per_rank=8)
    assert pf.epoch(10) is None and pf.epoch(10) == 5 and pf.search.bounds == [5]
    assert P.solve_batch_sharded(lambda g, lo: np.full(len(g), 3), np.zeros((4, 1, 1))).tolist() == [3, 3, 3, 3]


class NativeScripted(ScriptedSearch):
    """Stand-in for a search on an engine with a communicator: run() already exchanged the bound."""

    def global_best(self):
        return 7 if self.epoch >= 1 else None


def test_native_exchange_path_reads_global_best_only():
    pf = P.Portfolio(lambda off, n: NativeScripted(off, n, [9, 9, 9]), chains_per_rank=4, native=True)
    assert pf.epoch(10) is None and pf.epoch(10) == 7 and pf.epoch(10) == 7
    assert pf.search.bounds == []          # the engine owns the bound: nothing is pushed from the host
