"""Smallest run that touches every CUDA kernel of libtss once — meant to be run under
`compute-sanitizer --tool memcheck` (after the same command exited 0 without it).  Not a test: it only checks the
calls succeed; parity lives in tests/test_gpu.py."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import timberborn_support_solver_b200 as T  # noqa: E402

eng = T.Engine(0)
rng = np.random.default_rng(0)
one = T.PLATFORMS_DEFAULT[:1]

# kernel (a): small (8/16/32-bit rows), tiled sites, platforms (+ layers)
for w, h in [(7, 5), (16, 16), (21, 16), (32, 32), (40, 33)]:
    g = T.WorldGrid((rng.random((h, w)) < 0.8).astype(np.uint8))
    eng.eval_sites(g, (rng.random((37, h, w)) < 0.1).astype(np.uint8))
g = T.WorldGrid(np.ones((16, 16), np.uint8))
lay = [T.Platform(0, 0, T.PlatformDef(5, 5)), T.Platform(4, 4, T.PlatformDef(3, 3)), T.Platform(14, 14, T.PlatformDef(1, 4), True)]
eng.eval_platforms(g, [lay, []])
eng.validate(g, lay)
enc = T.Encoding.encode(T.PLATFORMS_DEFAULT, g)
a = eng.layout_to_assignment(enc, T.PlatformLayout(lay[:1]))

# kernel (c): check + propagate
cnf = enc.with_limits(T.PlatformLimits.new_unweighted({T.PlatformDef(1, 1): 3}))
dev = eng.upload_cnf(cnf)
full = np.full((5, cnf.n_vars + 1), 2, np.uint8)
full[:, : len(a)] = a
dev.propagate(full)
dev.check(rng.integers(0, 3, (33, cnf.n_vars + 1)).astype(np.uint8))

# kernel (b): half-warp kernel (<= 16 rows), full-warp kernel, batch, multi-platform, window decomposition
for grid in (np.ones((16, 16), np.uint8), (rng.random((32, 32)) < 0.7).astype(np.uint8)):
    s = eng.search(T.WorldGrid(grid), seed=1, n_chains=37)
    s.run(300, 0)
    s.run(300, 0)
    s.best_count()
    s.best_layout()
    s.read_chains()
    s.close()
eng.solve_batch(np.stack([(rng.random((32, 32)) < 0.7).astype(np.uint8) for _ in range(5)]), seed=1, steps=300, want_layouts=True)
eng.solve_batch(np.stack([(rng.random((9, 12)) < 0.7).astype(np.uint8) for _ in range(5)]), seed=1, steps=300, chains_per_terrain=16)
eng.solve_upper_bound(T.WorldGrid(np.ones((6, 5), np.uint8)), T.PLATFORMS_DEFAULT, card_limit=None, seed=1, max_steps=300)
s = eng.search(T.WorldGrid((rng.random((70, 50)) < 0.7).astype(np.uint8)), seed=1, n_chains=4)
s.run(200, 0)
s.run(200, 0)
s.best_layout()
s.close()
eng.solve_upper_bound(g, one, card_limit=15, seed=1)
eng.measure_peaks()
print("sanitize_smoke ok", eng.stats())
