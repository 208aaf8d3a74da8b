#!/usr/bin/env python3
"""Effect of the initial-guess order and of the multigrid smoothing parameters on the step, one set-up.
Usage: python tools/exp_guess.py [mesh]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from problems import TaylorGreen, make_mesh, make_solver  # noqa: E402
from oasisx_b200 import _lib as L  # noqa: E402
import bench  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
DT, NU = bench.DT, bench.NU
tg = TaylorGreen(NU, 3)
s = make_solver(make_mesh(3, N), 2, tg, DT, solver_options=bench.KRYLOV)
ctx = s._ctx
tg.t_u, tg.t_p = 0.0, -DT / 2


def steps(n, tag, every=True):
    rows = []
    for _ in range(n):
        tg.t_u += DT
        tg.t_p += DT
        s.solve(DT, NU, max_iter=1)
        st = ctx.stats()
        rows.append((st.ms_step, st.ms_assemble_first, st.ms_tentative, st.ms_pressure, st.ms_update,
                     max(st.its_tentative), st.its_pressure, max(st.its_update), st.res0_tentative, st.res0_pressure, st.res0_update))
        if every:
            r = rows[-1]
            print(f"  {tag} step {len(rows):2d}: {r[0]:7.3f} ms | first {r[1]:6.3f} tent {r[2]:6.3f} pres {r[3]:6.3f} upd {r[4]:6.3f} | its {r[5]}/{r[6]}/{r[7]} "
                  f"| res0 {r[8]:.1e} {r[9]:.1e} {r[10]:.1e}", flush=True)
    m = np.median(np.array(rows)[-5:], axis=0)
    print(f"{tag:36s} median(last 5) step {m[0]:7.3f} ms | first {m[1]:6.3f} tent {m[2]:6.3f} pres {m[3]:6.3f} upd {m[4]:6.3f} | its {int(m[5])}/{int(m[6])}/{int(m[7])}",
          flush=True)


for which in (L.SOLVER_TENTATIVE, L.SOLVER_SCALAR):
    ctx.set_solver_option(which, "b200_guess", "extrapolate")
steps(10, "guess=extrapolate")
for which in (L.SOLVER_TENTATIVE, L.SOLVER_SCALAR):
    ctx.set_solver_option(which, "b200_guess", "extrapolate2")
steps(16, "guess=extrapolate2")
for pre, post, om in ((1, 1, 0.8), (2, 2, 0.8), (1, 1, 0.857)):
    ctx.pressure_mg_configure(pre, post, 16, om)
    steps(6, f"extrapolate2, mg V({pre},{post}) om={om}", every=False)
