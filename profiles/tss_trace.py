"""TSS_TRACE=1 stage timing of tss_solve_instance / tss_witness_for_cnf (wall microseconds per stage, last warm repeat):
    python profiles/tss_trace.py ex1:default ex2:1x1 ..."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import timberborn_support_solver_b200 as T
from conftest import golden, rows_to_grid
exe = os.path.join(os.path.dirname(T.__file__), "tss_repl")
grids = {k: rows_to_grid(v["grid"]) for k, v in golden("fixtures").items()}
grids["readme"] = rows_to_grid(golden("readme_layouts")["terrain"])
with tempfile.TemporaryDirectory() as tmp:
    for spec in sys.argv[1:] or ["ex1:default"]:
        name, pset = spec.split(":")
        path = os.path.join(tmp, name + ".toml")
        open(path, "w").write(T.WorldGrid(grids[name]).to_toml())
        reps = 6
        r = subprocess.run([exe, path, "--platforms", pset, "--seed", "3", "--quiet", "--repeat", str(reps)], capture_output=True, text=True, env=dict(os.environ, TSS_TRACE="1"))
        lines = [ln for ln in r.stderr.splitlines() if ln.startswith("[tss trace]")]
        per = len(lines) // reps
        print(spec, "\n".join(lines[-per:]), sep="\n")
