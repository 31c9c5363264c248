#!/usr/bin/env python
"""A DIMACS command-line SAT solver for the tests of tools/tss_repl.cpp (`--exact CMD`): stands where Glucose stands in the
reference.  It is the oracle's CDCL (TEST INFRASTRUCTURE) behind the usual competition output format."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.oracle as O  # noqa: E402

lits, offsets, n_vars = [], [0], 0
for line in open(sys.argv[1]):
    t = line.split()
    if not t or t[0] == "c":
        continue
    if t[0] == "p":
        n_vars = int(t[2])
        continue
    lits += [int(x) for x in t[:-1]]
    offsets.append(len(lits))
a = np.full(n_vars + 1, 2, np.uint8)
la, oa = np.ascontiguousarray(lits, np.int32), np.ascontiguousarray(offsets, np.uint32)
r = O.lib().tsso_solve_csr(O._p(la), O._p(oa, C.c_uint32), len(offsets) - 1, n_vars, O._p(a, C.c_uint8), C.c_long(-1))
if r == 10:
    print("s SATISFIABLE")
    print("v " + " ".join(str(v if a[v] == 1 else -v) for v in range(1, n_vars + 1)) + " 0")
elif r == 20:
    print("s UNSATISFIABLE")
else:
    print("s UNKNOWN")
