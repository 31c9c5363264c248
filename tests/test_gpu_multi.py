"""Two-GPU test of the engine's native exchange (comm.cu): skipped on a single-GPU box.  Ranks are processes; the only
thing the host moves between them is the 128-byte NCCL id (through a file here)."""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, tmp, out):
    import timberborn_support_solver_b200 as T
    eng = T.Engine(rank)
    id_path = os.path.join(tmp, "nccl_id")
    if rank == 0:
        with open(id_path + ".tmp", "wb") as f:
            f.write(eng.comm_unique_id())
        os.replace(id_path + ".tmp", id_path)
    while not os.path.exists(id_path):
        time.sleep(0.01)
    eng.comm_init(open(id_path, "rb").read(), rank, world)
    assert eng.comm_world() == world
    grid = T.WorldGrid(np.ones((16, 16), np.uint8))
    s = eng.search(grid, seed=5, n_chains=64, chain_offset=rank * 64)
    hist = []
    for _ in range(6):
        s.run(300, 0)
        hist.append((s.best_count(), s.global_best()))
    chains = s.read_chains()
    s.close()
    # window-decomposed portfolio on a larger grid (C4 shape): after each phase every rank holds, per window, the best result of all ranks
    big = T.WorldGrid.synthetic(96, 80, 1, 3)
    s = eng.search(big, seed=9, n_chains=8, chain_offset=rank * 100000)
    counts = []
    for _ in range(5):
        s.run(1200, 0)
        counts.append(s.global_best())
    lay = sorted((p.x, p.y) for p in s.best_layout().platforms().values())     # validated by kernel (a) inside the engine
    s.close()
    single = None
    if rank == 0:       # the same phases on ONE GPU (an engine without a communicator, rank 0's seeds): what the second GPU has to beat
        solo = T.Engine(0)
        s1 = solo.search(big, seed=9, n_chains=8, chain_offset=0)
        for _ in range(5):
            s1.run(1200, 0)
        single = s1.best_count()
        s1.close()
        solo.close()
    out.put((rank, hist, int(chains["best"].min()), counts, lay, single))
    eng.close()


def test_native_nccl_portfolio_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, str(tmp_path), out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, h0, b0, c0, lay0, single), (_, h1, b1, c1, lay1, _) = res
    assert c0 == c1 and all(b <= a for a, b in zip(c0, c0[1:])) and len(lay0) == c0[-1]   # same counts on every rank, never increasing
    assert lay0 == lay1                                              # ... and the very same layout (assembled window by window from the winners)
    assert c0[-1] <= single, (c0, single)                            # two GPUs combine per window: at equal phases never worse than one
    # ... and it is exactly what the scalar replay of the two-rank portfolio assembles (oracle.lns_model: per window the rank with the
    # fewest core supports, lowest rank on ties), phase by phase
    import oracle.oracle as O
    import timberborn_support_solver_b200 as T
    want = O.lns_model(T.WorldGrid.synthetic(96, 80, 1, 3).data, 8, 5, 1200, seed=9, flat=True, threads=4, ranks=[(0, 20), (100000, 20)])
    assert [c for _, c in want] == c0
    assert lay0 == sorted((int(x), int(y)) for y, x in zip(*np.nonzero(want[-1][0])))
    print("C4-shaped portfolio, 5 phases: two GPUs", c0, "one GPU", single)
    for (l0, g0), (l1, g1) in zip(h0, h1):
        assert g0 == g1                                             # every rank sees the same global bound after each epoch
        locals_ = [x for x in (l0, l1) if x is not None]
        if locals_:
            assert g0 == min(locals_)                               # ... and it is the min over the ranks' best counts
    assert h0[-1][1] == 15 == min(b0, b1)                           # the portfolio reaches the proven optimum of rect 16x16
