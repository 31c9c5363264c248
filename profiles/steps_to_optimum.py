"""How many steps does the fastest of SM x 16 half-warp chains need to reach the proven optimum, as a function of the epoch
length E (the chains share their best count between epochs)?"""
import json, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import timberborn_support_solver_b200 as T
eng = T.Engine(0)
n = eng.device_info()["sm_count"] * 16
fx = json.load(open("tests/golden/fixtures.json"))
ex2 = np.array([[1 if (i < len(r) and r[i] == "X") else 0 for i in range(21)] for r in fx["ex2"]["grid"]], np.uint8)
for name, grid, opt in (("rect16", np.ones((16, 16), np.uint8), 15), ("ex2", ex2, 14)):
    for E in (4, 8, 16, 32, 64):
        need = []
        for seed in range(40):
            s = eng.search(T.WorldGrid(grid), seed=1000 + seed, n_chains=n, kernel=T.KERNEL_HALF_WARP)
            s.set_bound(opt + 1)
            steps = 0
            while s.best_count() is None and steps < 512:
                s.run(E, opt)
                steps += E
            need.append(steps)
            s.close()
        need = np.array(need)
        print(name, "E", E, "steps to optimum: min", need.min(), "median", np.median(need), "p90", np.percentile(need, 90), "max", need.max(), flush=True)
