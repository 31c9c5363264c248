"""Kernel (c) check: contiguous clause slices per block (default) vs the strided slices of r1 (TSS_CNF_STRIDED=1), interleaved
in one process on the bench's instance (6505 clauses x 131072 assignments)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import timberborn_support_solver_b200 as T
eng = T.Engine(0)
grid = T.WorldGrid(np.ones((16, 16), np.uint8))
enc = T.Encoding.encode(T.PLATFORMS_DEFAULT[:1], grid)
cnf = enc.with_limits(T.PlatformLimits.new_unweighted({T.PlatformDef(1, 1): 15}))
dev = eng.upload_cnf(cnf)
res, wit = eng.solve_upper_bound(grid, card_limit=15, seed=300)
full = np.full((1, cnf.n_vars + 1), 2, np.uint8)
base = eng.layout_to_assignment(enc, wit)
full[0, : len(base)] = base
prop, conflict, rounds = dev.propagate(full)
prop[prop == 2] = 0
a = np.repeat(prop, 131072, axis=0)
modes = [("strided", "1", "0", ""), ("chunked", "0", "0", "")]
out = {m[0]: [] for m in modes}
for rep in range(8):
    for name, strided, swap, gy in modes:
        os.environ["TSS_CNF_STRIDED"] = strided
        nf, _ = dev.check(a)
        assert nf.sum() == 0
        if rep >= 2:
            out[name].append(eng.stats()["device_ms"])
for mode, ts in out.items():
    print(f"{mode}: median {np.median(ts):.4f} ms, min {min(ts):.4f}, max {max(ts):.4f}")
