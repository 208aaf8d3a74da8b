#!/usr/bin/env python3
"""k_spmm with its fused dot products accumulated per warp (tuning "spmm_warp_dots") against per thread: timings of the
SpMM with one / two fused dot products at 96^3 (b2_bench_kernel 4 / 5), whole-step time, and the fields of three steps on
a small box either way.  NEGATIVE RESULT: the tuning key exists only in commit 9937232; kept as the
script that produced profiles/r02_spmm_warp_dots_96cube.txt.  Usage: python tools/experiments/exp_warp_dots.py [mesh]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from problems import TaylorGreenRot, make_mesh, make_solver, relerr, vscale  # noqa: E402
import bench  # noqa: E402

dt, nu = 0.005, 0.01
fields = []
for wd in (0, 1):
    tg = TaylorGreenRot(nu)
    s = make_solver(make_mesh(3, 12), 2, tg, dt, solver_options=bench.KRYLOV)
    s._ctx.set_tuning("spmm_warp_dots", wd)
    tg.t_u, tg.t_p = 0.0, -dt / 2
    its = []
    for _ in range(4):
        tg.t_u += dt
        tg.t_p += dt
        s.solve(dt, nu, max_iter=1)
        st = s._ctx.stats()
        its.append((tuple(st.its_tentative), st.its_pressure, tuple(st.its_update)))
    fields.append(([s._u[i].x.array_ro().copy() for i in range(3)], s._p.x.array_ro().copy(), its))
(u0, p0, i0), (u1, p1, i1) = fields
print("12^3, 4 steps: max rel diff u", max(relerr(u1[i], u0[i], vscale(u0)) for i in range(3)), "p", relerr(p1, p0), "its equal", i0 == i1, i0[-1], flush=True)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
tg = TaylorGreenRot(nu)
s = make_solver(make_mesh(3, N), 2, tg, dt, solver_options=bench.KRYLOV)
ctx = s._ctx
tg.t_u, tg.t_p = 0.0, -dt / 2
for _ in range(12):
    tg.t_u += dt
    tg.t_p += dt
    s.solve(dt, nu, max_iter=1)
for rep in range(2):
    for wd in (0, 1):
        ctx.set_tuning("spmm_warp_dots", wd)
        t = [ctx.bench_kernel(k, 20)[0] for k in (3, 4, 5)]
        ms = []
        for _ in range(6):
            tg.t_u += dt
            tg.t_p += dt
            s.solve(dt, nu, max_iter=1)
            st = ctx.stats()
            ms.append((st.ms_step, st.ms_tentative, st.ms_update, max(st.its_tentative), max(st.its_update)))
        m = np.median(np.array(ms), axis=0)
        print(f"warp_dots={wd}: spmm k3 {t[0]:.4f} k4 {t[1]:.4f} k5 {t[2]:.4f} ms | step {m[0]:.3f} tentative {m[1]:.3f} update {m[2]:.3f} ms (its {int(m[3])}/{int(m[4])})", flush=True)
