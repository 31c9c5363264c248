"""Per-step latency of the warp and half-warp kernels on a lightly loaded device (rect 16x16, bound 16, never done):
what a one-shot solve pays per step in latency mode."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import timberborn_support_solver_b200 as T
eng = T.Engine(0)
grid = T.WorldGrid(np.ones((16, 16), np.uint8))
sm = eng.device_info()["sm_count"]
for per_sm in (4, 8, 16, 32):
    row = []
    for kernel in (T.KERNEL_WARP, T.KERNEL_HALF_WARP):
        s = eng.search(grid, seed=1, n_chains=sm * per_sm, kernel=kernel)
        s.set_bound(16)
        s.run(500, -1); s.best_count()
        s.run(4000, -1); s.best_count()
        row.append(eng.stats()["device_ms"] / 4000 * 1e3)
        s.close()
    print(per_sm, "chains/SM: warp %.2f us/step  half-warp %.2f us/step" % tuple(row), flush=True)
