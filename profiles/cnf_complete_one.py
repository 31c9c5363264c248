"""Witness completion (csrc/cnf.cu cnf_complete_kernel) on test/ex2.toml with the default-8 platform set and the limit of the last
SAT iteration: the SLS layout of 4 platforms completed into a model of the 25 K clauses.  Used under ncu (-k regex:cnf_complete) and
on its own for the device time (batch kernels beside it: tss_cnf_propagate + tss_cnf_check, the r2 path it replaced)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import timberborn_support_solver_b200 as T
eng = T.Engine(0)
fx = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fixtures.json")))
rows = fx["ex2"]["grid"]
w = max(len(r) for r in rows)
g = T.WorldGrid(np.array([[1 if (i < len(r) and r[i] == "X") else 0 for i in range(w)] for r in rows], np.uint8))
enc = T.Encoding.encode(T.PLATFORMS_DEFAULT, g)
res, lay = eng.solve_upper_bound(g, T.PLATFORMS_DEFAULT, card_limit=None, seed=3, max_steps=200000)
n = lay.platform_count()
cnf = enc.with_limits(T.PlatformLimits.new_unweighted({T.PlatformDef(1, 1): n}))
dev = eng.upload_cnf(cnf)
base = eng.layout_to_assignment(enc, lay)
full = np.full(cnf.n_vars + 1, 2, np.uint8)
full[: len(base)] = base
for rep in range(5):
    t0 = time.perf_counter()
    a, conflict, nf = dev.complete(full)
    wall = (time.perf_counter() - t0) * 1e6
    fused_us = eng.stats()["device_ms"] * 1e3
t0 = time.perf_counter()
prop, c2, rounds = dev.propagate(full[None, :])
prop[prop == 2] = 0
nf2, _ = dev.check(prop)
batch_wall = (time.perf_counter() - t0) * 1e6
assert conflict < 0 and nf == 0 and c2[0] < 0 and nf2[0] == 0 and np.array_equal(a[1:], prop[0][1:])
print(json.dumps({"instance": "test/ex2.toml, default-8 set, at most %d platforms" % n, "vars": cnf.n_vars, "clauses": cnf.n_clauses,
                  "fused_kernel_us": fused_us, "fused_call_wall_us": wall, "batch_path_wall_us": batch_wall, "sync_rounds": rounds}))
