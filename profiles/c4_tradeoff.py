"""configs[3] (256x256, p = 0.7): count after a fixed number of chain steps, split into more / fewer phases (a phase = one window
offset), and the effect of chains per window.  One B200."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import timberborn_support_solver_b200 as T
eng = T.Engine(0)
g = T.WorldGrid.synthetic(256, 256, 1, 0)
for phases, steps, chains in [(8, 15000, 0), (16, 7500, 0), (32, 3750, 0), (64, 1875, 0), (128, 940, 0), (34, 3750, 0), (16, 7500, 16), (16, 7500, 32), (64, 1875, 16), (64, 3750, 0), (128, 1875, 0)]:
    s = eng.search(g, seed=1, n_chains=chains)
    t0 = time.perf_counter()
    for _ in range(phases + 1):
        s.run(steps, 0)
    c = s.best_count()
    ms = (time.perf_counter() - t0) * 1e3
    print(f"phases {phases + 1:4d} x {steps:6d} steps, chains/window {s.n_chains // 81:3d}: count {c}  ({ms:.0f} ms)")
    s.close()
