"""Builds timberborn_support_solver_b200/libtss.so (the C-ABI library of include/tss.h) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libtss.so")
SOURCES = ["engine.cu", "eval.cu", "eval_thread.cu", "cnf.cu", "sls.cu", "sls_h16.cu", "sls_t16.cu", "sls_multi.cu", "lns.cu", "greedy.cu", "comm.cu", "peaks.cu", "capi_host.cpp", "host_model.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
              "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas", "-shared", "-cudart", "static", "-ldl"]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "tss.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", OUT, *[os.path.join(CSRC, s) for s in SOURCES]]
    subprocess.check_call(cmd, cwd=CSRC)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
