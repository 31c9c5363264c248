import os, subprocess, sys, tempfile
sys.path.insert(0, "/root/repo")
import numpy as np
import timberborn_support_solver_b200 as T
exe = os.path.join(os.path.dirname(T.__file__), "tss_repl")
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "rect24.toml")
open(path, "w").write(T.WorldGrid(np.ones((24, 24), np.uint8)).to_toml())
for args in (["--gui"], [], ["--platforms", "1x1"]):
    r = subprocess.run([exe, path, "--seed", "3", "--quiet", "--repeat", "3"] + args, capture_output=True, text=True, timeout=600)
    print(args, r.stdout.strip().splitlines()[-1] if r.stdout else r.stderr[-300:])
