"""One `load; solve` of the C++ driver (tss_repl) on a named instance: python profiles/repl_one.py ex2 1x1 [extra tss_repl args].
Used under ncu for the launch list of a whole bound-tightening loop (which kernels a solve launches, and their shares)."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import timberborn_support_solver_b200 as T
from conftest import golden, rows_to_grid
grids = {k: rows_to_grid(v["grid"]) for k, v in golden("fixtures").items()}
grids["readme"] = rows_to_grid(golden("readme_layouts")["terrain"])
name, pset = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, name + ".toml")
open(path, "w").write(T.WorldGrid(grids[name]).to_toml())
exe = os.path.join(os.path.dirname(T.__file__), "tss_repl")
os.execv(exe, [exe, path, "--platforms", pset, "--seed", "3"] + sys.argv[3:])
