"""C3 (BASELINE.json configs[2]): README 21x16 terrain / test/ex2.toml, 1x1 supports, one-shot solve at the proven optimum 14."""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import timberborn_support_solver_b200 as T
eng = T.Engine(0)
fx = json.load(open("tests/golden/fixtures.json"))
readme = json.load(open("tests/golden/readme_layouts.json"))
def grid_from_rows(rows):
    w = max(len(r) for r in rows)
    return np.array([[1 if (i < len(r) and r[i] == "X") else 0 for i in range(w)] for r in rows], np.uint8)
grids = {"ex2": grid_from_rows(fx["ex2"]["grid"])}
g = grids["ex2"].copy(); g[8:10, 8:10] = 1          # README terrain = ex2 with (8,8),(9,8),(8,9),(9,9) set to ceiling
grids["readme"] = g
for name, grid in grids.items():
    ts, steps = [], []
    for seed in range(24):
        s0 = eng.stats()
        t0 = time.perf_counter()
        res, lay = eng.solve_upper_bound(T.WorldGrid(grid), card_limit=14, seed=500 + seed)
        ts.append((time.perf_counter() - t0) * 1e3)
        s1 = eng.stats()
        steps.append((s1["sls_steps"] - s0["sls_steps"]) / max(1, 148 * 16))
        assert res == T.SAT and lay.platform_count() == 14, (res, lay and lay.platform_count())
    ts = np.array(ts[4:])
    print(name, int(grid.sum()), "tiles: median", round(float(np.median(ts)), 3), "ms  p90", round(float(np.percentile(ts, 90)), 3), "ms  steps/chain median", float(np.median(steps[4:])))
