"""Randomized parity campaign for the paths added late in round 2 (one B200):
  * window decomposition (csrc/lns.cu) against oracle.lns_model (flat port in WINDOW mode): random grid sizes 33..110, densities, seeds,
    chains per window, phase lengths — the global layout after every phase must be the same set of supports;
  * terrain batch (tss_solve_batch) against the flat port: random terrain sizes <= 32x32, densities, step counts — same count per terrain;
  * fused witness completion (tss_cnf_complete) against the oracle's unit propagation on random clause sets.
python profiles/parity_campaign_r2b.py [n_lns] [n_batch] [n_cnf]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import timberborn_support_solver_b200 as T
import oracle.oracle as O
from conftest import synth_terrain
n_lns, n_batch, n_cnf = (int(a) for a in (sys.argv[1:4] + ["60", "40", "300"][len(sys.argv) - 1:]))
eng = T.Engine(0)
rng = np.random.default_rng(20261018)
t0 = time.time()
bad = 0
for case in range(n_lns):
    w, h = int(rng.integers(33, 111)), int(rng.integers(33, 111))
    if rng.random() < 0.3:
        w = int(rng.integers(8, 33)) if rng.random() < 0.5 else w
        h = int(rng.integers(33, 111)) if w <= 32 else h
    dens = float(rng.choice([0.35, 0.55, 0.7, 0.85, 1.0]))
    grid = synth_terrain(w, h, seed=int(rng.integers(1, 1000)), t=int(rng.integers(0, 1000)), density_q24=int(dens * (1 << 24)))
    if grid.sum() == 0:
        continue
    seeds, phases, steps, seed = int(rng.choice([4, 8, 12])), int(rng.integers(2, 6)), int(rng.choice([300, 900, 2000])), int(rng.integers(0, 1 << 30))
    want = O.lns_model(grid, seeds, phases, steps, seed=seed, flat=True, threads=8)
    s = eng.search(T.WorldGrid(grid), seed=seed, n_chains=seeds)
    for p, (S, count) in enumerate(want):
        s.run(steps, 0)
        got = np.zeros((h, w), np.uint8)
        for pl in s.best_layout().platforms().values():
            got[pl.y, pl.x] = 1
        if s.best_count() != count or not np.array_equal(got, S):
            bad += 1
            print("LNS MISMATCH", case, w, h, dens, seeds, phases, steps, seed, "phase", p, s.best_count(), count)
            break
    s.close()
print(f"window decomposition: {n_lns} random cases, {bad} mismatches ({time.time() - t0:.0f} s)")
t0 = time.time()
bad_b = n_terr = 0
for case in range(n_batch):
    w, h = int(rng.integers(6, 33)), int(rng.integers(6, 33))
    dens = float(rng.choice([0.4, 0.7, 0.9, 1.0]))
    n = int(rng.integers(8, 65))
    steps = int(rng.choice([500, 1024, 1500, 2600]))
    seed = int(rng.integers(0, 1 << 30))
    base = int(rng.integers(0, 100000))
    grids = np.stack([synth_terrain(w, h, seed=7, t=base + t, density_q24=int(dens * (1 << 24))) for t in range(n)])
    counts = eng.solve_batch(grids, seed=seed, steps=steps)
    per = 8 if h <= 16 else 4          # chains per terrain of tss_solve_batch (two chains per warp when the grid has <= 16 rows)
    epochs = [(min(1024, steps - d), 1 << 20, 0) for d in range(0, steps, 1024)]
    for t in range(n):
        if grids[t].sum() == 0:
            continue
        r = O.sls_flat(grids[t], per, epochs, seed=seed, chain_offset=per * t, want_layouts=False)
        n_terr += 1
        if int(r["best"].min()) != counts[t]:
            bad_b += 1
            print("BATCH MISMATCH", case, w, h, dens, steps, seed, t, int(r["best"].min()), counts[t])
print(f"terrain batch: {n_terr} terrains in {n_batch} random batches, {bad_b} mismatches ({time.time() - t0:.0f} s)")
t0 = time.time()
bad_c = 0
for case in range(n_cnf):
    n_vars = int(rng.choice([5, 30, 200, 1500, 9000, 40000]))
    n_cl = int(n_vars * rng.uniform(0.5, 4.0)) + 1
    hidden = rng.integers(0, 2, n_vars + 1).astype(np.uint8) if rng.random() < 0.6 else None
    clauses = []
    order = rng.permutation(n_vars) + 1
    for i in range(n_vars // 2):
        a, b = int(order[i % n_vars]), int(order[(i + 1) % n_vars])
        clauses.append((-a if rng.random() < 0.8 else a, b if rng.random() < 0.8 else -b))
    for _ in range(n_cl):
        k = int(rng.choice([1, 2, 2, 3, 3, 3, 4, 5, 7, 12]))
        vs = rng.choice(n_vars, size=min(k, n_vars), replace=False) + 1
        clauses.append(tuple(int(v) if rng.random() < 0.5 else -int(v) for v in vs))
    if hidden is not None:
        clauses = [c if any(hidden[abs(l)] == (1 if l > 0 else 0) for l in c) else c[:-1] + (-c[-1],) for c in clauses]
    clauses = [clauses[i] for i in rng.permutation(len(clauses))]
    lits = np.array([l for c in clauses for l in c], np.int32)
    offs = np.zeros(len(clauses) + 1, np.uint32)
    offs[1:] = np.cumsum([len(c) for c in clauses])
    a = np.full(n_vars + 1, 2, np.uint8)
    dec = rng.random(n_vars + 1) < rng.choice([0.0, 0.1, 0.4, 0.8])
    a[dec] = hidden[dec] if hidden is not None else rng.integers(0, 2, int(dec.sum()))
    a[0] = 2
    want, wc, _ = O.propagate_csr(lits, offs, n_vars, a)
    got, conflict, nf = eng.upload_cnf(T.Cnf(n_vars, lits, offs)).complete(a)
    ok = (conflict >= 0) == (wc >= 0)
    if ok and wc < 0:
        want[want == 2] = 0
        ok = np.array_equal(got[1:], want[1:])
    if not ok:
        bad_c += 1
        print("CNF MISMATCH", case, n_vars, n_cl, conflict, wc)
print(f"witness completion: {n_cnf} random clause sets, {bad_c} mismatches ({time.time() - t0:.0f} s)")
sys.exit(1 if bad or bad_b or bad_c else 0)
