#!/usr/bin/env python3
"""CUDA-graphed pressure iterations on/off.  Usage: python tools/exp_graph.py [mesh ...]"""
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
    tg = TaylorGreen(NU, 3)
    s = make_solver(make_mesh(3, N), 2, tg, DT, solver_options=bench.KRYLOV)
    ctx = s._ctx
    tg.t_u, tg.t_p = 0.0, -DT / 2

    def steps(n):
        r = []
        for _ in range(n):
            tg.t_u += DT
            tg.t_p += DT
            s.solve(DT, NU, max_iter=1)
            st = ctx.stats()
            r.append((st.ms_step, st.ms_pressure, st.its_pressure))
        return np.median(np.array(r), axis=0)

    steps(30)
    for g in (1, 0, 1, 0):
        ctx.set_tuning("graphs", g)
        m = steps(12)
        print(f"N={N} graphs={g}: step {m[0]:.3f} ms, pressure stage {m[1]:.3f} ms, its {int(m[2])}", flush=True)
    del s, ctx
