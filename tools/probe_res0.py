#!/usr/bin/env python3
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from problems import TaylorGreen, make_mesh, make_solver
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dt, nu = 0.005, 0.01
for guess in ("nonzero", "extrapolate", "extrapolate+atol"):
    tg = TaylorGreen(nu, 3)
    opts = {k: {"ksp_type": t, "pc_type": "jacobi", "ksp_rtol": 1e-10} for k, t in (("tentative", "bcgs"), ("pressure", "cg"), ("scalar", "cg"))}
    opts["pressure"]["pc_type"] = "mg"
    for k in opts:
        if guess != "none":
            opts[k]["ksp_initial_guess_nonzero"] = True
        if guess.startswith("extrapolate") and k != "pressure":
            opts[k]["b200_guess"] = "extrapolate"
        if guess.endswith("atol") and k != "pressure":
            opts[k]["ksp_atol"] = 1e-16
    s = make_solver(make_mesh(3, N), 2, tg, dt, solver_options=opts)
    tg.t_u, tg.t_p = 0.0, -dt / 2
    for n in range(10):
        tg.t_u += dt; tg.t_p += dt
        s.solve(dt, nu, max_iter=1)
        st = s.stats()
        if n in (2, 5, 9):
            print(f"{guess:12s} step {n}: res0 t/p/u {st.res0_tentative:.2e} {st.res0_pressure:.2e} {st.res0_update:.2e}  its tentative {list(st.its_tentative)} pressure {st.its_pressure} update {list(st.its_update)}", flush=True)
