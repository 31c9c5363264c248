"""BASELINE.json configs[0]: `load test/exN.toml; solve` — the REPL's bound-tightening loop (crates/repl/src/main.rs:280-366)
end to end, CPU only (oracle CDCL as the Glucose stand-in, 1 thread) against the same loop with the GPU engine as the SAT
side and the CDCL solver called once, for the final UNSAT proof.  Wall clock, host buffers."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import oracle.oracle as O
import timberborn_support_solver_b200 as T
from oracle.oracle import _p

def exact(cnf):
    a = np.full(cnf.n_vars + 1, 2, np.uint8)
    lits, offs = np.ascontiguousarray(cnf.lits, np.int32), np.ascontiguousarray(cnf.offsets, np.uint32)
    r = O.lib().tsso_solve_csr(_p(lits), _p(offs, C.c_uint32), cnf.n_clauses, cnf.n_vars, _p(a, C.c_uint8), C.c_long(-1))
    return {10: T.SAT, 20: T.UNSAT}.get(r, T.INTERRUPTED), a

eng = T.Engine(0)
fx = json.load(open("tests/golden/fixtures.json"))
out = {}
for name in ("ex1", "ex3", "ex2"):
    rows = fx[name]["grid"]
    w = max(len(r) for r in rows)
    grid = np.array([[1 if (i < len(r) and r[i] == "X") else 0 for i in range(w)] for r in rows], np.uint8)
    for label, defs, odefs in (("default-8", T.PLATFORMS_DEFAULT, O.PLATFORMS_DEFAULT), ("1x1", T.PLATFORMS_DEFAULT[:1], O.PLATFORMS_1X1)):
        t0 = time.perf_counter()
        r = O.solver_loop(grid, odefs, conflict_budget=20_000_000)
        cpu_s = time.perf_counter() - t0
        cpu_counts = [s.get("count") for s in r["steps"] if s["result"] == 10]
        g = T.WorldGrid(grid)
        ts = []
        for rep in range(3):
            t0 = time.perf_counter()
            enc = T.Encoding.encode(defs, g)
            res = T.solver_loop(T.Project(T.World(g)), enc, T.PlatformLimits(), eng, exact_solver=exact, seed=rep)
            ts.append(time.perf_counter() - t0)
        gpu_steps = [s for s in res["steps"] if s["source"] == "gpu"]
        last = res["steps"][-1]
        line = dict(cpu_loop_s=round(cpu_s, 3), cpu_solves=len(r["steps"]), cpu_optimum=min(cpu_counts), gpu_loop_s=round(min(ts), 4), gpu_solves=len(gpu_steps),
                    gpu_best=res["best"].platform_count(), proved=res["proved_optimal"], final=f'{last["result"]} by {last["source"]}')
        out[f"{name} {label}"] = line
        print(name, label, line, flush=True)
json.dump(out, open("gpurun_out/repl_loop_timing.json", "w"), indent=1)
