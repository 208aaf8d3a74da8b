"""assemble_first (k_first_cells) against the cell schedule (slab thickness of the class interleaving; 0 = mesh order) and kernel time at N^3.
    python tools/exp_first.py [N] [bricks...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from problems import make_mesh, make_solver  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
sizes = [int(a) for a in sys.argv[2:]] or [4, 0, 1, 2, 8, 16]  # slab thickness; 0 = mesh order
tg = bench.make_field("taylor-green-rot")
s = make_solver(make_mesh(3, N), 2, tg, bench.DT, solver_options=bench.krylov_for("taylor-green-rot"))
ctx = s._ctx
for _ in range(2):
    tg.t_u += bench.DT
    tg.t_p += bench.DT
    s.solve(bench.DT, bench.NU, max_iter=1)
for b in sizes:
    ctx.set_tuning("first_order", 1 if b > 0 else 0)
    if b > 0:
        ctx.set_tuning("first_slab", b)
    ms, nbytes = ctx.bench_kernel(1, 10)
    print(f"first_slab={b}: {ms:.3f} ms  algorithmic {nbytes / 1e9:.2f} GB -> {nbytes / ms / 1e6:.0f} GB/s  plan {ctx.first_plan_info()}", flush=True)
