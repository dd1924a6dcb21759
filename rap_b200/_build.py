"""In-tree build of the sm_100a shared library (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librap_b200.so")
SOURCES = ["api.cu", "gemm.cu", "fmha.cu", "elementwise.cu", "procrustes.cu", "spinnet.cu", "preproc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-rdc=true", "-Xptxas", "-v"]


def _newer(a: str, b: str) -> bool:
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "rap_b200.h"))
    objs, procs = [], []
    for src in SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(objdir, src + ".o")
        objs.append(o)
        if force or _newer(s, o) or any(_newer(h, o) for h in headers):
            log = open(o + ".log", "w")
            procs.append((src, subprocess.Popen([nvcc, *NVCC_FLAGS, "-c", s, "-o", o], stdout=log, stderr=subprocess.STDOUT), o + ".log"))
    failed = False
    for src, p, log in procs:
        if p.wait() != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src} ---\n{open(log).read()}\n")
        elif verbose:
            sys.stderr.write(open(log).read())
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or not os.path.exists(LIB):
        subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
