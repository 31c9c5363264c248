"""The C++ driver (timberborn_support_solver_b200/tss_repl) on the named instances: one whole `load; solve` to the PROVEN optimum
over the C ABI, `--repeat 21` (cold first run, warm median of the rest).  Writes the driver's summary lines as JSON.

    python profiles/tss_repl_timing.py > profiles/r2_tss_repl_timing.json
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import timberborn_support_solver_b200 as T   # noqa: E402
from conftest import golden, rows_to_grid      # noqa: E402

exe = os.path.join(os.path.dirname(T.__file__), "tss_repl")
grids = {k: rows_to_grid(v["grid"]) for k, v in golden("fixtures").items()}
grids["readme"] = rows_to_grid(golden("readme_layouts")["terrain"])
out = {}
with tempfile.TemporaryDirectory() as tmp:
    for name in ("ex1", "ex3", "ex2", "readme"):
        path = os.path.join(tmp, name + ".toml")
        with open(path, "w") as f:
            f.write(T.WorldGrid(grids[name]).to_toml())
        for pset in ("default", "1x1"):
            r = subprocess.run([exe, path, "--platforms", pset, "--seed", "3", "--quiet", "--repeat", "21", "--phases"], capture_output=True, text=True, timeout=600)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("# ")]
            out[f"{name} {pset}"] = lines[0] if r.returncode == 0 else r.stderr[-300:]
            if r.returncode == 0 and len(lines) > 1:
                out[f"{name} {pset} phases"] = lines[1]
print(json.dumps(out, indent=1))
