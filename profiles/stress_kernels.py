"""One-off stress: thread-per-chain vs half-warp kernel, 20000 chains x 150000 steps on four terrains; every chain state must agree."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import timberborn_support_solver_b200 as T
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from conftest import synth_terrain
eng = T.Engine(0)
for name, grid in [("rect16", np.ones((16, 16), np.uint8)), ("r21x16", synth_terrain(21, 16, seed=3, t=9)), ("r26x13", synth_terrain(26, 13, seed=4, t=1, density_q24=int(0.5 * (1 << 24)))),
                   ("r16x16-dense", synth_terrain(16, 16, seed=5, t=2, density_q24=int(0.9 * (1 << 24))))]:
    res = {}
    for kernel in (T.KERNEL_THREAD, T.KERNEL_HALF_WARP):
        s = eng.search(T.WorldGrid(grid), seed=77, n_chains=20000, chain_offset=123456, kernel=kernel)
        t0 = time.perf_counter()
        for steps in (50000, 100000):
            s.run(steps, 0)
        st = s.read_chains()
        res[kernel] = (st, time.perf_counter() - t0, s.best_count())
        s.close()
    a, b = res[T.KERNEL_THREAD], res[T.KERNEL_HALF_WARP]
    ok = all(np.array_equal(a[0][k], b[0][k]) for k in ("S", "bestS", "k", "best", "step", "scored"))
    print(name, "agree" if ok else "MISMATCH", "best", a[2], b[2], f"thread {a[1]:.2f} s, half-warp {b[1]:.2f} s", flush=True)
    assert ok
