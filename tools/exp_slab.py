"""Per-rank kernel efficiency of an 8-way strong-scaling run, measured on ONE GPU: the 96 x 96 x 12 slab one rank of
eight holds at 96^3.  SpMM (P2xP2, 3 right-hand sides) against the persistent-grid size, assemble_first, Ap SpMV.
    python tools/exp_slab.py [nz]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from oasisx_b200 import mesh as bmesh  # noqa: E402
from problems import make_solver  # noqa: E402

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 12
tg = bench.make_field("taylor-green-rot")
msh = bmesh.create_box(None, [[-1.0, -1.0, -1.0], [1.0, 1.0, -1.0 + 2.0 * nz / 96]], [96, 96, nz])
s = make_solver(msh, 2, tg, bench.DT, solver_options=bench.krylov_for("taylor-green-rot"))
ctx = s._ctx
for _ in range(3):
    tg.t_u += bench.DT
    tg.t_p += bench.DT
    s.solve(bench.DT, bench.NU, max_iter=1)
peak = bench.measured_peaks()[0]
for bps, mins in ((8, 0), (8, 1)):
    ctx.set_tuning("spmm_blocks_per_sm", bps)
    ctx.set_tuning("spmm_min_slices", mins)
    for u in (8,):
        ctx.set_tuning("spmm_unroll", u)
        ms, nb = ctx.bench_kernel(3, 50)
        print(f"slab nz={nz}: spmm blocks/SM={bps} min_slices={mins} unroll={u}: {ms * 1e3:.1f} us  {nb / ms / 1e6:.0f} GB/s = {nb / ms / 1e6 / peak:.2f} of peak", flush=True)
ctx.set_tuning("spmm_blocks_per_sm", 8)
ctx.set_tuning("spmm_min_slices", 1)
ctx.set_tuning("spmm_unroll", 8)
ms, nb = ctx.bench_kernel(1, 20)
print(f"assemble_first: {ms * 1e3:.1f} us ({nb / ms / 1e6:.0f} GB/s)")
ms, nb = ctx.bench_kernel(2, 100)
print(f"Ap spmv: {ms * 1e3:.1f} us ({nb / ms / 1e6:.0f} GB/s)")
st = s.stats()
print(f"step: {st.ms_step:.3f} ms  stages {st.ms_assemble_first:.3f} {st.ms_tentative:.3f} {st.ms_pressure:.3f} {st.ms_update:.3f}  its {list(st.its_tentative)}/{st.its_pressure}/{list(st.its_update)}")
