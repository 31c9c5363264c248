"""configs[3]: chains per window against time and count (one warm-up phase, then 16 phases; python profiles/c4_chains.py [steps])."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import timberborn_support_solver_b200 as T
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 7500
eng = T.Engine(0)
g = T.WorldGrid.synthetic(256, 256, 1, 0)
for seed in (1, 2):
    for chains in (4, 8, 12, 16, 24, 56):
        s = eng.search(g, seed=seed, n_chains=chains)
        s.run(steps, 0); s.best_count()
        t0 = time.perf_counter()
        for _ in range(16):
            s.run(steps, 0)
        c = s.best_count()
        ms = (time.perf_counter() - t0) * 1e3
        print(f"seed {seed} chains/window {s.n_chains // 81:3d}: count {c}  ({ms:.0f} ms for 16 phases x {steps} steps)")
        s.close()
