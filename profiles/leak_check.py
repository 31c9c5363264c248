import sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import timberborn_support_solver_b200 as T
from conftest import golden, rows_to_grid
g = rows_to_grid(golden("fixtures")["ex2"]["grid"])
eng = T.Engine(0)
grid = T.WorldGrid(g)
enc = T.Encoding.encode(T.PLATFORMS_DEFAULT, grid)
def used():
    torch.cuda.synchronize()
    free, total = torch.cuda.mem_get_info()
    return (total - free) / 2**20
one = T.PlatformDef(1, 1)
res = None
for rep in range(3):
    m0 = used()
    t0 = time.time()
    for i in range(1500):
        cnf = enc.with_limits(T.PlatformLimits.new_unweighted({one: 4 + (i % 3)}))
        dev = eng.upload_cnf(cnf)
        if i % 50 == 0:
            res = T.solver_loop(T.Project(T.World(grid)), enc, T.PlatformLimits(), eng, exact_solver=None, seed=i)
        del dev
    print(f"round {rep}: {used() - m0:+.1f} MiB after 1500 uploads + 30 loops ({time.time() - t0:.1f} s), device memory in use {used():.0f} MiB")
print("last loop:", res["best"].platform_count(), res["proved_optimal"])
