"""Krylov iterations and initial residuals (relative to the tolerance's reference norm) per step over the start-up
transient of the 96^3 benchmark.   python tools/probe_transient.py [N] [steps] [workload]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from problems import make_mesh, make_solver  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
wl = sys.argv[3] if len(sys.argv) > 3 else "taylor-green-rot"
tg = bench.make_field(wl)
s = make_solver(make_mesh(3, N), 2, tg, bench.DT, solver_options=bench.krylov_for(wl))
for n in range(steps):
    tg.t_u += bench.DT
    tg.t_p += bench.DT
    s.solve(bench.DT, bench.NU, max_iter=1)
    st = s.stats()
    print(f"step {n + 1:3d}: its {list(st.its_tentative)} / {st.its_pressure} / {list(st.its_update)}   res0 {st.res0_tentative:.2e} {st.res0_pressure:.2e} "
          f"{st.res0_update:.2e}   ms {st.ms_step:.2f} = {st.ms_assemble_first:.2f} + {st.ms_tentative:.2f} + {st.ms_pressure:.2f} + {st.ms_update:.2f}", flush=True)
