"""Fractional bounds whose tableau does not fit the cluster's shared memory: the cooperative grid simplex (one CTA per SM, one grid barrier
per pivot) against the one-CTA kernel (TSS_LP_SINGLE_CTA=1), same instances: default-8 sets on the 21x16 terrains, 1x1 supports on a
24x24 rectangle (the GUI's default grid) and on a 32x32 random terrain, the GUI's weights on 24x24."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import timberborn_support_solver_b200 as T
eng = T.Engine(0)
fx = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fixtures.json")))
rows = fx["ex2"]["grid"]
w = max(len(r) for r in rows)
ex2 = np.array([[1 if (i < len(r) and r[i] == "X") else 0 for i in range(w)] for r in rows], np.uint8)
GUI = {T.PlatformDef(1, 1): 5, T.PlatformDef(1, 2): 1, T.PlatformDef(1, 3): 1, T.PlatformDef(1, 4): 1, T.PlatformDef(1, 5): 1, T.PlatformDef(1, 6): 1, T.PlatformDef(3, 3): 2, T.PlatformDef(5, 5): 4}
cases = [("ex2 default-8", ex2, T.PLATFORMS_DEFAULT, None), ("rect 24x24 1x1", np.ones((24, 24), np.uint8), T.PLATFORMS_DEFAULT[:1], None),
         ("random 32x32 1x1", T.WorldGrid.synthetic(32, 32, 1, 0).data, T.PLATFORMS_DEFAULT[:1], None),
         ("rect 24x24 default-8, GUI weights", np.ones((24, 24), np.uint8), T.PLATFORMS_DEFAULT, GUI)]
for name, grid, defs, wts in cases:
    out = {}
    for mode in ("grid", "one CTA"):
        os.environ["TSS_LP_SINGLE_CTA"] = "1" if mode == "one CTA" else "0"
        if mode == "one CTA" and "GUI" in name:
            continue      # (minutes on one SM)
        if mode == "grid":
            eng.lower_bound_lp(T.WorldGrid(grid), defs, weights=wts)      # warm-up: scratch buffers of this size
        t0 = time.perf_counter()
        r = eng.lower_bound_lp(T.WorldGrid(grid), defs, weights=wts)
        out[mode] = dict(bound=r["bound"], pivots=r["pivots"], optimal=r["optimal"], constraints=r["constraints"], device_ms=round(eng.stats()["device_ms"], 2), wall_ms=round((time.perf_counter() - t0) * 1e3, 1))
    print(name, json.dumps(out))
