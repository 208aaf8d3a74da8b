#!/usr/bin/env python3
"""How predictable are the tentative velocity u* and the correction delta = u - u* from step to step?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from problems import TaylorGreen, make_mesh, make_solver

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dt, nu = 0.005, 0.01
tg = TaylorGreen(nu, 3)
opts = {k: {"ksp_type": t, "pc_type": "jacobi", "ksp_rtol": 1e-12} for k, t in (("tentative", "bcgs"), ("pressure", "cg"), ("scalar", "cg"))}
opts["pressure"]["pc_type"] = "mg"
s = make_solver(make_mesh(3, N), 2, tg, dt, solver_options=opts)
tg.t_u, tg.t_p = 0.0, -dt / 2
us, ustars, deltas, dps = [], [], [], []
nrm = lambda a: float(np.linalg.norm(a))
for n in range(8):
    tg.t_u += dt; tg.t_p += dt
    s._ps.x.array[:] = s._p.x.array_ro()
    [[bc.update_bc() for bc in b] for b in s._bcs_u]
    s.assemble_first(dt, nu); s.velocity_tentative_assemble(); s.velocity_tentative_solve()
    ustar = np.concatenate([s._u[i].x.array_ro().copy() for i in range(3)])
    s.pressure_assemble(dt); s.pressure_solve(nu)
    dp = s._dp.x.array_ro().copy()
    s.velocity_update(dt)
    u = np.concatenate([s._u[i].x.array_ro().copy() for i in range(3)])
    for i in range(3):
        s._u2[i].x.array[:] = s._u1[i].x.array_ro(); s._u1[i].x.array[:] = s._u[i].x.array_ro()
    s._p.x.array[:] = s._ps.x.array_ro()
    us.append(u); ustars.append(ustar); deltas.append(u - ustar); dps.append(dp)
    if n >= 2:
        d, d1, d2 = deltas[-1], deltas[-2], deltas[-3]
        print(f"step {n}: |delta|/|u| {nrm(d)/nrm(u):.2e}  |d-d1|/|d| {nrm(d-d1)/nrm(d):.2e}  |d-(2d1-d2)|/|d| {nrm(d-2*d1+d2)/nrm(d):.2e}  "
              f"|u*-u1*|/|u| {nrm(ustar-ustars[-2])/nrm(u):.2e}  |u*-(2u1-u2)|/|u| {nrm(ustar-(2*us[-2]-us[-3]))/nrm(u):.2e}  "
              f"|u*-(2u1-u2-d1)|/|u| {nrm(ustar-(2*us[-2]-us[-3]-d1))/nrm(u):.2e}  |u*-(2u1*-u2*)|/|u| {nrm(ustar-(2*ustars[-2]-ustars[-3]))/nrm(u):.2e}  "
              f"|dp-dp1|/|dp| {nrm(dp-dps[-2])/nrm(dp):.2e}")
