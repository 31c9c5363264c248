"""One-off randomized parity campaign: every SLS kernel variant that fits vs the scalar CPU model on random terrains,
densities, seeds, noise levels, bounds and warm starts.  Prints a summary; any mismatch raises."""
import os, sys, time
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import oracle.oracle as O
import timberborn_support_solver_b200 as T
from conftest import synth_terrain

eng = T.Engine(0)
rng = np.random.default_rng(2026)
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
t16_only = len(sys.argv) > 2 and sys.argv[2] == "t16"     # only grids the thread-per-chain kernel takes (<= 16 rows x 26 columns)
t0 = time.time()
checked = 0
for case in range(n_cases):
    w, h = (int(rng.integers(1, 27)), int(rng.integers(1, 17))) if t16_only else (int(rng.integers(1, 33)), int(rng.integers(1, 33)))
    dens = float(rng.choice([0.3, 0.5, 0.7, 0.85, 1.0]))
    grid = synth_terrain(w, h, seed=int(rng.integers(1, 1000)), t=int(rng.integers(0, 1000)), density_q24=int(dens * (1 << 24)))
    if grid.sum() == 0:
        continue
    seed, offset = int(rng.integers(0, 1 << 30)), int(rng.integers(0, 100000))
    noise = int(rng.choice([0, 10, 20, 50]))
    n_chains = int(rng.integers(3, 12))
    bound = (1 << 20) if rng.random() < 0.6 else int(grid.sum() // int(rng.integers(6, 14)) + 1)
    epochs = [(int(rng.integers(1, 700)), bound, int(rng.choice([-1, 0]))) for _ in range(int(rng.integers(1, 4)))]
    warm = None
    if rng.random() < 0.4:
        warm = np.zeros((n_chains, 32, 32), np.uint8)
        warm[:, :h, :w] = rng.random((n_chains, h, w)) < rng.choice([0.05, 0.3, 1.0])
    want = O.sls_model(grid, n_chains, epochs, seed=seed, chain_offset=offset, noise_pct=noise, share_bound=True, init_S=warm)
    flat = O.sls_flat(grid, n_chains, epochs, seed=seed, chain_offset=offset, noise_pct=noise, share_bound=True, init_S=warm, threads=2)
    assert all(np.array_equal(flat[key], want[key]) for key in ("S", "bestS", "k", "best", "step", "scored", "steps")), ("flat port", case)
    kernels = [T.KERNEL_WARP] + ([T.KERNEL_HALF_WARP] if h <= 16 else []) + ([T.KERNEL_THREAD] if h <= 16 and w <= 26 else [])
    for kernel in kernels:
        s = eng.search(T.WorldGrid(grid), seed=seed, n_chains=n_chains, chain_offset=offset, noise_pct=noise, kernel=kernel)
        flips0 = eng.stats()["sls_flips"]
        if warm is not None:
            s.write_chains((warm.astype(np.uint32) << np.arange(32, dtype=np.uint32)).sum(2, dtype=np.uint32))
        if bound < (1 << 20):
            s.set_bound(bound)
        for steps, _, target in epochs:
            s.run(steps, target)
        got = s.read_chains()
        assert eng.stats()["sls_flips"] - flips0 == int(flat["flips"].sum()), (case, kernel, "flips")
        unpack = lambda r: ((r[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(np.uint8)
        for key in ("k", "best", "step", "scored"):
            assert np.array_equal(got[key], want[key]), (case, kernel, key, w, h, dens, seed, epochs)
        assert np.array_equal(unpack(got["S"]), want["S"]) and np.array_equal(unpack(got["bestS"]), want["bestS"]), (case, kernel, w, h)
        s.close()
        checked += 1
print(f"{checked} kernel runs over {n_cases} random cases agree with the CPU model and the flat CPU port bit for bit, flips included ({time.time() - t0:.1f} s)")
