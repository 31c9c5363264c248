"""One-off randomized parity campaign for the placement search (platform sets beyond {1x1}) vs its scalar CPU model:
random terrains, platform sets, weights, seeds, bounds.  Also kernel (a) on the winners: the oracle's validate() accepts
every best layout with exactly the reported count."""
import os, sys, time
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import oracle.oracle as O
import timberborn_support_solver_b200 as T
from conftest import synth_terrain

eng = T.Engine(0)
rng = np.random.default_rng(7)
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
t0 = time.time()
done = validated = 0
for case in range(n_cases):
    w, h = int(rng.integers(3, 33)), int(rng.integers(3, 33))
    grid = synth_terrain(w, h, seed=int(rng.integers(1, 1000)), t=int(rng.integers(0, 1000)), density_q24=int(float(rng.choice([0.5, 0.7, 0.9, 1.0])) * (1 << 24)))
    if grid.sum() == 0:
        continue
    pool = [(1, 2), (1, 3), (2, 2), (1, 4), (2, 3), (3, 3), (1, 6), (2, 5), (4, 4), (5, 5), (3, 6), (6, 6)]
    defs = [T.PlatformDef(1, 1)] + [T.PlatformDef(*pool[i]) for i in rng.choice(len(pool), size=int(rng.integers(1, 6)), replace=False)]
    weights = None
    if rng.random() < 0.35:
        weights = {d: int(rng.integers(1, 9)) for d in defs if rng.random() < 0.7 or d == defs[0]}
    seed, offset, n_chains = int(rng.integers(0, 1 << 30)), int(rng.integers(0, 5000)), int(rng.integers(2, 7))
    epochs = [(int(rng.integers(5, 160)), 1 << 20, int(rng.choice([-1, 0]))) for _ in range(int(rng.integers(1, 4)))]
    s = eng.search(T.WorldGrid(grid), defs, seed=seed, n_chains=n_chains, chain_offset=offset)
    try:
        if weights:
            s.set_weights(weights)
    except T.TssError:
        s.close()
        continue
    for steps, _, target in epochs:
        s.run(steps, target)
    got = s.read_placements()
    kd = got["key_dims"]
    canon = lambda a, b: (min(a, b), max(a, b))
    costs = [1] * len(kd) if not weights else [sum(v for d, v in weights.items() if canon(d.width, d.height)[0] <= canon(a, b)[0] and canon(d.width, d.height)[1] <= canon(a, b)[1]) for a, b in kd]
    want = O.slsm_model(grid, kd, costs, n_chains, epochs, seed=seed, chain_offset=offset, share_bound=True)
    for key in ("k", "best", "best_k", "step"):
        assert np.array_equal(got[key], want[key]), (case, key, w, h, [tuple((d.width, d.height)) for d in defs], weights)
    assert np.array_equal(got["items"], want["items"]) and np.array_equal(got["best_items"], want["best_items"]), (case, w, h)
    for c in np.nonzero(got["best"] < (1 << 20))[0][:2]:
        plats = []
        for code in got["best_items"][c][: got["best_k"][c]]:
            a, b = kd[int(code) >> 10]
            plats.append((int(code) & 31, (int(code) >> 5) & 31, min(a, b), max(a, b), int(a > b)))
        v = O.validate(grid, plats)
        assert v.is_valid and sum(costs[int(code) >> 10] for code in got["best_items"][c][: got["best_k"][c]]) == got["best"][c]
        validated += 1
    s.close()
    done += 1
print(f"{done} random placement-search cases agree with the CPU model bit for bit; {validated} best layouts pass the oracle's validate() ({time.time() - t0:.1f} s)")
