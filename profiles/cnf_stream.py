"""Kernel (c) on the bench's instance: the 16x16 encoder CNF + totalizer of the at-most-15 bound (6505 clauses), 131072
assignments (the SLS witness completed by unit propagation, every 64th with one support removed); used plain and under ncu."""
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
a[::64, int(enc.vars().plat_var[[p.y * 16 + p.x for p in wit.platforms().values()][0], 0])] = 0
for _ in range(3):
    nf, _ = dev.check(a)
ms = eng.stats()["device_ms"]
assert (nf[::64] > 0).all() and nf.reshape(-1, 64)[:, 1:].sum() == 0
print(f"cnf check: {cnf.n_clauses} clauses x {len(a)} assignments in {ms:.3f} ms = {cnf.n_clauses * len(a) / ms / 1e9:.2f} T clause evaluations/s, "
      f"{len(cnf.lits) * (len(a) // 32) * 4 / ms / 1e6:.0f} GB/s of plane reads, {rounds} propagation rounds")
