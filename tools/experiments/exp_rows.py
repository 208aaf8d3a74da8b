#!/usr/bin/env python3
"""Row-wise fused assemble_first against the scatter + combine pair at full size: time and step effect.
Usage: python tools/exp_rows.py [mesh]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from problems import TaylorGreen, make_mesh, make_solver  # noqa: E402
import bench  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
DT, NU = bench.DT, bench.NU
tg = TaylorGreen(NU, 3)
s = make_solver(make_mesh(3, N), 2, tg, DT, solver_options=bench.KRYLOV)
ctx = s._ctx
tg.t_u, tg.t_p = 0.0, -DT / 2
for _ in range(4):
    tg.t_u += DT
    tg.t_p += DT
    s.solve(DT, NU, max_iter=1)
ctx.set_tuning("assemble_rows", 0)
for comb in (0, 1, 2, 3, 4, 0):
    ctx.set_tuning("combine", comb)
    ms, nb = ctx.bench_kernel(1, 20)
    print(f"combine variant {comb}: assemble_first {ms:.3f} ms", flush=True)
ctx.set_tuning("combine", int(os.environ.get("EXP_COMBINE", "0")))
for rows in (1, 0):
    ctx.set_tuning("assemble_rows", rows)
    ms, nb = ctx.bench_kernel(1, 10)
    print(f"assemble_rows={rows}: assemble_first {ms:.3f} ms", flush=True)
    b = [s._b_first[i].x.array_ro().copy() for i in range(3)]
    print("   |b_first| per component:", [float(np.linalg.norm(v)) for v in b], flush=True)
for rows in (1, 0):
    ctx.set_tuning("assemble_rows", rows)
    r = []
    for _ in range(8):
        tg.t_u += DT
        tg.t_p += DT
        s.solve(DT, NU, max_iter=1)
        st = ctx.stats()
        r.append((st.ms_step, st.ms_assemble_first, st.ms_tentative, st.ms_pressure, st.ms_update, max(st.its_tentative), st.its_pressure, max(st.its_update)))
    m = np.median(np.array(r), axis=0)
    print(f"assemble_rows={rows}: step {m[0]:.3f} ms | first {m[1]:.3f} tent {m[2]:.3f} pres {m[3]:.3f} upd {m[4]:.3f} | its {int(m[5])}/{int(m[6])}/{int(m[7])}", flush=True)
