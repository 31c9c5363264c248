"""time-to-optimal (rect 16x16, card_limit = 15, one-shot solve from host buffers) vs chains per SM of the latency
configuration: TSS_EXPERIMENT_HALF_WARP_CHAINS_PER_SM=N python profiles/tto_sweep.py"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import timberborn_support_solver_b200 as T
eng = T.Engine(0)
grid = T.WorldGrid(np.ones((16, 16), np.uint8))
ts, steps = [], []
for seed in range(44):
    s0 = eng.stats()
    t0 = time.perf_counter()
    res, lay = eng.solve_upper_bound(grid, card_limit=15, seed=1000 + seed)
    ts.append((time.perf_counter() - t0) * 1e3)
    s1 = eng.stats()
    steps.append((s1["sls_steps"] - s0["sls_steps"], s1["kernel_launches"] - s0["kernel_launches"]))
    assert res == T.SAT and lay.platform_count() == 15
ts = np.array(ts[4:])
print(os.environ.get("TSS_EXPERIMENT_HALF_WARP_CHAINS_PER_SM", "default"), f"median {np.median(ts):.3f} ms  p10 {np.percentile(ts, 10):.3f}  p90 {np.percentile(ts, 90):.3f}  launches/call {np.median([l for _, l in steps[4:]])}")
