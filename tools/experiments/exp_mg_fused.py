#!/usr/bin/env python3
"""Fused multigrid kernels (tuning "mg_fused") on/off -- same fields (bitwise for the
first+residual pair), same iteration counts, pressure-stage time (measured at 48^3: 1.27 ms fused against 1.01 ms, fields equal to 1e-13).  Usage: python tools/exp_mg_fused.py [mesh ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from problems import TaylorGreen, make_mesh, make_solver  # noqa: E402
import bench  # noqa: E402

DT, NU = bench.DT, bench.NU
for N in [int(a) for a in sys.argv[1:]] or [96, 48]:
    sols = {}
    for fused in (0, 1):
        tg = TaylorGreen(NU, 3)
        s = make_solver(make_mesh(3, N), 2, tg, DT, solver_options=bench.KRYLOV)
        ctx = s._ctx
        ctx.set_tuning("mg_fused", fused)
        tg.t_u, tg.t_p = 0.0, -DT / 2
        r = []
        for _ in range(40):
            tg.t_u += DT
            tg.t_p += DT
            s.solve(DT, NU, max_iter=1)
            st = ctx.stats()
            r.append((st.ms_step, st.ms_pressure, st.its_pressure))
        m = np.median(np.array(r[-10:]), axis=0)
        sols[fused] = (s._p.x.array_ro().copy(), [s._u[i].x.array_ro().copy() for i in range(3)])
        print(f"N={N} mg_fused={fused}: step {m[0]:.3f} ms, pressure stage {m[1]:.3f} ms, its {int(m[2])}", flush=True)
        del s, ctx
    dp = np.abs(sols[0][0] - sols[1][0]).max() / np.abs(sols[0][0]).max()
    du = max(np.abs(a - b).max() for a, b in zip(sols[0][1], sols[1][1]))
    print(f"N={N}: max rel difference of p after 40 steps {dp:.2e}, max abs difference of u {du:.2e}", flush=True)
