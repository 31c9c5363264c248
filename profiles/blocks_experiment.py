import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import timberborn_support_solver_b200 as T
import oracle.oracle as O
from conftest import synth_terrain
eng = T.Engine(0)
for (w, h, dens, t) in [(96, 72, 0.85, 1), (128, 128, 0.7, 0), (256, 256, 0.7, 0)]:
    grid = synth_terrain(w, h, seed=4 if w == 96 else 1, t=t, density_q24=int(dens * (1 << 24)))
    g = T.WorldGrid(grid)
    t0 = time.perf_counter()
    s = eng.search(g, T.PLATFORMS_DEFAULT, seed=5)
    for _ in range(6):
        s.run(2000, 0)
    lay = s.best_layout(); s.close()
    t_big = time.perf_counter() - t0
    # block decomposition: every 32x32 block solved on its own with the placement search
    t0 = time.perf_counter()
    plats = []
    for by in range(0, h, 32):
        for bx in range(0, w, 32):
            blk = np.ascontiguousarray(grid[by:by + 32, bx:bx + 32])
            if blk.sum() == 0: continue
            res, bl = eng.solve_upper_bound(T.WorldGrid(blk), T.PLATFORMS_DEFAULT, seed=5, max_steps=3000)
            assert res == T.SAT
            for p in bl.platforms().values():
                plats.append((p.x + bx, p.y + by, p.definition.width, p.definition.height, int(p.rotated)))
    t_blk = time.perf_counter() - t0
    v = O.validate(grid, plats)
    print(f"{w}x{h} p={dens}: engine (1x1 LNS + merge / greedy) {lay.platform_count()} platforms in {t_big*1e3:.0f} ms; independent 32x32 blocks {len(plats)} platforms in {t_blk*1e3:.0f} ms, valid={v.is_valid}")
