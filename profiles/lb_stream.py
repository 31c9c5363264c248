"""The packing lower bound (tss_lower_bound, csrc/lb.cu) on the named instances: bound, time, witness size.  Used plain and under ncu."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import timberborn_support_solver_b200 as T
eng = T.Engine(0)
fx = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fixtures.json")))
proofs = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "proofs.json")))
def grid_of(name):
    if name == "rect16":
        return np.ones((16, 16), np.uint8)
    rows = fx["ex2" if name == "readme" else name]["grid"]
    w = max(len(r) for r in rows)
    g = np.array([[1 if (i < len(r) and r[i] == "X") else 0 for i in range(w)] for r in rows], np.uint8)
    if name == "readme":
        g[8:10, 8:10] = 1
    return g
out = {}
for name in ("ex1", "ex3", "ex2", "readme", "rect16"):
    for label, defs in (("1x1", T.PLATFORMS_DEFAULT[:1]), ("default8", T.PLATFORMS_DEFAULT)):
        g = T.WorldGrid(grid_of(name))
        eng.lower_bound(g, defs, seed=1)
        ts = []
        for rep in range(5):
            t0 = time.perf_counter()
            tiles = eng.lower_bound(g, defs, seed=1 + rep)
            ts.append((time.perf_counter() - t0) * 1e3)
        opt = proofs.get(f"{name}/{label}", {}).get("optimum")
        dev_ms = eng.stats()["device_ms"]
        eng.lower_bound_lp(g, defs)
        tl = []
        for rep in range(3):
            t0 = time.perf_counter()
            lp = eng.lower_bound_lp(g, defs)
            tl.append((time.perf_counter() - t0) * 1e3)
        out[f"{name}/{label}"] = {"packing_bound": len(tiles), "fractional_bound": lp["bound"], "proven_optimum": opt, "packing_ms_wall": round(float(np.median(ts)), 3),
                                  "packing_ms_device": round(dev_ms, 3), "lp_ms_wall": round(float(np.median(tl)), 3), "lp_ms_device": round(eng.stats()["device_ms"], 3),
                                  "lp_value": round(lp["total"] / lp["max_load"], 4), "lp_pivots": lp["pivots"], "lp_constraints": lp["constraints"], "lp_optimal": lp["optimal"]}
        print(name, label, out[f"{name}/{label}"], flush=True)
rng = np.random.default_rng(0)
g32 = T.WorldGrid.synthetic(32, 32, 1, 0)
tiles = eng.lower_bound(g32, T.PLATFORMS_DEFAULT[:1], seed=1)
pk_ms = eng.stats()["device_ms"]
lp = eng.lower_bound_lp(g32, T.PLATFORMS_DEFAULT[:1])
out["C5 terrain 0 (32x32 p=0.7)/1x1"] = {"packing_bound": len(tiles), "packing_ms_device": round(pk_ms, 3), "fractional_bound": lp["bound"], "lp_value": round(lp["total"] / lp["max_load"], 4),
                                         "lp_ms_device": round(eng.stats()["device_ms"], 3), "lp_pivots": lp["pivots"], "lp_optimal": lp["optimal"],
                                         "sls_count": int(eng.solve_batch(g32.data[None], seed=1, steps=20000, chains_per_terrain=32)[0])}
print(out["C5 terrain 0 (32x32 p=0.7)/1x1"])
json.dump(out, open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "r2_lower_bounds.json"), "w"), indent=1)
