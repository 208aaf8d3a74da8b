#!/usr/bin/env python3
"""A/B experiments on one GPU with ONE set-up of the benchmark problem: SpMM variants (run-compressed columns,
register bounds), multigrid coarse solve / smoothing parameters, and their effect on the whole IPCS step.
Usage: python tools/exp_step.py [mesh] [steps]"""
import json
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
K = int(sys.argv[2]) if len(sys.argv) > 2 else 6
DT, NU = bench.DT, bench.NU
tg = TaylorGreen(NU, 3)
msh = make_mesh(3, N)
s = make_solver(msh, 2, tg, DT, solver_options=bench.KRYLOV)
ctx = s._ctx
slots, runs = ctx.pattern_sell(L.PAT_VV)
nnz = ctx.pattern_nnz(L.PAT_VV)
print(json.dumps({"mesh": N, "nnz_VV": nnz, "slots_VV": slots, "slice_cols": slots // 32, "runs": runs,
                  "run_fraction": runs / (slots // 32)}), flush=True)
tg.t_u, tg.t_p = 0.0, -DT / 2


def steps(n):
    out = []
    for _ in range(n):
        tg.t_u += DT
        tg.t_p += DT
        s.solve(DT, NU, max_iter=1)
        st = ctx.stats()
        out.append((st.ms_step, st.ms_assemble_first, st.ms_tentative, st.ms_pressure, st.ms_update,
                    max(st.its_tentative), st.its_pressure, max(st.its_update)))
    return np.array(out)


def report(tag, a):
    m = np.median(a, axis=0)
    print(f"{tag:42s} step {m[0]:7.3f} ms | first {m[1]:6.3f} tent {m[2]:6.3f} pres {m[3]:6.3f} upd {m[4]:6.3f} | its {int(m[5])}/{int(m[6])}/{int(m[7])}",
          flush=True)


steps(4)  # warm-up: histories for the extrapolated guesses
# ---- SpMM variants: equality with the plain kernel, then timing ------------------------------------------
n = s._nV_owned
x = np.random.default_rng(0).uniform(-1, 1, n)
ctx.set_tuning("spmm_comp", 0)
y0 = ctx.mat_mult(L.MAT_M, 0, x, n)
for comp in (1, 2, 3, 4):
    ctx.set_tuning("spmm_comp", comp)
    y = ctx.mat_mult(L.MAT_M, 0, x, n)
    print(f"spmm_comp={comp}: max |y - y_plain| = {np.abs(y - y0).max():.3e} (bitwise equal: {bool((y == y0).all())})", flush=True)
for comp in (0, 1, 2, 3, 4):
    ctx.set_tuning("spmm_comp", comp)
    for bps in ((8,) if comp in (0, 1, 2) else (6, 8)):
        ctx.set_tuning("spmm_blocks_per_sm", bps)
        ms, nb = ctx.bench_kernel(3, 20)
        ms_a, _ = ctx.bench_kernel(0, 20)
        print(f"spmm_comp={comp} blocks/SM={bps}: M {ms:.4f} ms ({nb / ms / 1e6:7.1f} GB/s by CSR bytes)  A {ms_a:.4f} ms", flush=True)
ctx.set_tuning("spmm_blocks_per_sm", 8)

# ---- whole step ---------------------------------------------------------------------------------------------
for comp in (0, 1, 2, 3, 4):
    ctx.set_tuning("spmm_comp", comp)
    report(f"step: spmm_comp={comp} mg_dense=1", steps(K))
best = int(os.environ.get("EXP_COMP", "1"))
ctx.set_tuning("spmm_comp", best)
ctx.set_tuning("mg_dense", 0)
report(f"step: spmm_comp={best} mg_dense=0", steps(K))
ctx.set_tuning("mg_dense", 1)
for pre, post, om in ((2, 2, 0.7), (2, 2, 0.8), (2, 2, 0.857), (1, 1, 0.8), (1, 1, 0.857), (2, 1, 0.8), (1, 2, 0.8), (3, 3, 0.8)):
    ctx.pressure_mg_configure(pre, post, 16, om)
    report(f"step: mg V({pre},{post}) omega={om}", steps(K))
