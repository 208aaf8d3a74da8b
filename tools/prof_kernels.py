#!/usr/bin/env python3
"""Small driver for ncu: sets up the 3D Taylor-Green problem and launches each hot kernel a few times.
Usage: python tools/prof_kernels.py [mesh] [kernel ids...]   (ids as in b2_bench_kernel)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from problems import TaylorGreen, make_mesh, make_solver  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kernels = [int(a) for a in sys.argv[2:]] or [0, 1, 2, 3]
dt, nu = 0.005, 0.01
tg = TaylorGreen(nu, 3)
opts = {k: {"ksp_type": t, "pc_type": "jacobi", "ksp_rtol": 1e-10}
        for k, t in (("tentative", "bcgs"), ("pressure", "cg"), ("scalar", "cg"))}
s = make_solver(make_mesh(3, N), 2, tg, dt, solver_options=opts)
tg.t_u, tg.t_p = dt, dt / 2
s.solve(dt, nu, max_iter=1)
for k in kernels:
    ms, nbytes = s._ctx.bench_kernel(k, 3)
    print(f"kernel {k}: {ms:.4f} ms  {nbytes / ms / 1e6:.1f} GB/s")
