"""One fractional-bound solve on test/ex2.toml (1x1 supports) and on the README terrain: the instances whose proof it closes.  Used under ncu."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import timberborn_support_solver_b200 as T
eng = T.Engine(0)
fx = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fixtures.json")))
rows = fx["ex2"]["grid"]
w = max(len(r) for r in rows)
g = np.array([[1 if (i < len(r) and r[i] == "X") else 0 for i in range(w)] for r in rows], np.uint8)
for name, grid in (("ex2", g), ("readme", None)):
    if grid is None:
        grid = g.copy()
        grid[8:10, 8:10] = 1
    r = eng.lower_bound_lp(T.WorldGrid(grid))
    print(name, {k: v for k, v in r.items() if k != "weights"}, "value", r["total"] / r["max_load"], "device ms", eng.stats()["device_ms"])
    tiles = eng.lower_bound(T.WorldGrid(grid), seed=1)
    print(name, "packing", len(tiles), "device ms", eng.stats()["device_ms"])
big = T.WorldGrid.synthetic(256, 256, 1, 0)
print("C4 packing", len(eng.lower_bound(big, seed=1)), "device ms", eng.stats()["device_ms"])
