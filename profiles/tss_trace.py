"""TSS_TRACE=1 stage timing of tss_solve_instance on ex1 (default-8): the last warm repeat's lines."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import timberborn_support_solver_b200 as T
from conftest import golden, rows_to_grid
exe = os.path.join(os.path.dirname(T.__file__), "tss_repl")
grids = {k: rows_to_grid(v["grid"]) for k, v in golden("fixtures").items()}
with tempfile.TemporaryDirectory() as tmp:
    for name in sys.argv[1:] or ["ex1"]:
        path = os.path.join(tmp, name + ".toml")
        open(path, "w").write(T.WorldGrid(grids[name]).to_toml())
        r = subprocess.run([exe, path, "--seed", "3", "--quiet", "--repeat", "6"], capture_output=True, text=True, env=dict(os.environ, TSS_TRACE="1"))
        lines = r.stderr.splitlines()
        per = len(lines) // 6
        print(name, "\n".join(lines[-per:]), sep="\n")
