"""C5 batch: mean best count vs (chains per terrain, steps per chain) at equal chain-steps per terrain."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import timberborn_support_solver_b200 as T
eng = T.Engine(0)
lib = T.load()
n = 4096
grids = np.zeros((n, 32, 32), np.uint8)
for t in range(n):
    lib.tss_world_synthetic(32, 32, 1, t, int(0.7 * (1 << 24)), grids[t].ctypes.data_as(C.POINTER(C.c_uint8)))
for cpt, steps in [(4, 2000), (8, 1000), (16, 500), (32, 250), (32, 400), (32, 600), (32, 1000), (32, 2000), (4, 4000), (8, 2000), (16, 1000)]:
    eng.solve_batch(grids[:256], seed=1, steps=steps, chains_per_terrain=cpt)
    t0 = time.perf_counter()
    c = eng.solve_batch(grids, seed=1, steps=steps, chains_per_terrain=cpt)
    dt = time.perf_counter() - t0
    print(f"chains {cpt:3d} steps {steps:5d}: mean count {c.mean():.3f}  {n / dt / 1e3:.1f} k terrains/s (wall)")
