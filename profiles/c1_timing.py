import json, os, time, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import timberborn_support_solver_b200 as T
eng = T.Engine(0)
fx = json.load(open("tests/golden/fixtures.json"))
for name, opt in [("ex1", 1), ("ex3", 1), ("ex2", 4)]:
    rows = fx[name]["grid"]
    g = T.WorldGrid.from_toml("[world]\ngrid = [\n" + "".join(f'    "{r}",\n' for r in rows) + "]\n")
    for rep in range(4):
        t0 = time.perf_counter()
        res, lay = eng.solve_upper_bound(g, T.PLATFORMS_DEFAULT, card_limit=opt, seed=rep)
        dt = (time.perf_counter() - t0) * 1e3
        st = eng.stats()
        print(name, rep, res, lay.platform_count() if lay else None, f"{dt:.2f} ms", "dev_ms", st["device_ms"], "launches", st["kernel_launches"], "steps", st["sls_steps"])
