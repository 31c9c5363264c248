"""One fractional bound through the grid-wide simplex: rect 24x24 with 1x1 supports (577 x 577 tableau, 2.7 MB).  Used under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import timberborn_support_solver_b200 as T
eng = T.Engine(0)
g = T.WorldGrid(np.ones((24, 24), np.uint8))
for _ in range(2):
    r = eng.lower_bound_lp(g)
    print({k: v for k, v in r.items() if k != "weights"}, "device ms", eng.stats()["device_ms"])
