"""Build oracle/libipcs_cpu.so (the C++/OpenMP CPU restatement; test + baseline infrastructure)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ipcs_cpu.cpp")
LIB = os.path.join(HERE, "libipcs_cpu.so")


def build(force: bool = False) -> str:
    deps = [SRC, os.path.join(HERE, "..", "oasisx_b200", "csrc", "ref_tables.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    # -march=x86-64-v3 (AVX2/FMA) rather than native: the .so is built here and runs on the GPU box's host
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-fopenmp", "-std=c++17", "-fPIC", "-shared", "-o", LIB, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building libipcs_cpu.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
