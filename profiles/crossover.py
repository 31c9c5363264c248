"""Auto-selection threshold: thread-per-chain vs half-warp kernel, device time for 20000 steps on rect 16x16 vs number of chains."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import timberborn_support_solver_b200 as T
eng = T.Engine(0)
grid = T.WorldGrid(np.ones((16, 16), np.uint8))
for n in (512, 1024, 2048, 4096, 8192, 16384):
    row = []
    for kernel in (T.KERNEL_THREAD, T.KERNEL_HALF_WARP):
        s = eng.search(grid, seed=1, n_chains=n, kernel=kernel)
        s.run(2000, 0); s.best_count()
        s.run(20000, 0); s.best_count()
        row.append(eng.stats()["device_ms"])
        s.close()
    print(n, f"thread {row[0]:.1f} ms  half-warp {row[1]:.1f} ms", flush=True)
