#!/usr/bin/env python3
"""Sweep SpMM kernels / launch shapes / slice schedules on the GPU (one set-up, every variant timed by b2_bench_kernel:
kernel 3 = M x (no reduction), 4 = with one fused dot product + row scale, 5 = with two).
Usage: python tools/sweep_spmm.py [mesh]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from problems import TaylorGreen, make_mesh, make_solver  # noqa: E402
from oasisx_b200 import _lib as L, fem  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
dt, nu = 0.005, 0.01
tg = TaylorGreen(nu, 3)
opts = {k: {"ksp_type": t, "pc_type": "jacobi", "ksp_rtol": 1e-10}
        for k, t in (("tentative", "bcgs"), ("pressure", "cg"), ("scalar", "cg"))}
msh = make_mesh(3, N)
s = make_solver(msh, 2, tg, dt, solver_options=opts, bricks=True)
ctx = s._ctx
tg.t_u, tg.t_p = dt, dt / 2
s.solve(dt, nu, max_iter=1)
Vs = s._Vi[0][0]
x = Vs.tabulate_dof_coordinates()
n = Vs.num_dofs
print("bricks:", s._brick_info, flush=True)


def run(tag):
    out = []
    for kern in (3, 4, 5):
        ms, nbytes = ctx.bench_kernel(kern, 10)
        out.append(f"k{kern} {ms:.4f} ms")
    print(f"{tag:60s} " + "  ".join(out), flush=True)


ctx.set_tuning("spmm_brick", 1)
run("brick kernel, first form (2 blocks x 16 warps, register-staged stream)")
ctx.set_tuning("spmm_brick", 2)
run("brick kernel, pipelined (1 block x 16 warps, TMA-fed rings)")
for mode, name in ((1, "no fill"), (2, "no stream")):
    ctx.set_tuning("spmm_brick_diag", mode)
    run(f"  pipelined, {name}")
ctx.set_tuning("spmm_brick_diag", 0)
ctx.set_tuning("spmm_brick", 3)
run("brick kernel, pipelined, 16-byte cp.async rings instead of TMA")
for mode, name in ((1, "no fill"), (2, "no stream")):
    ctx.set_tuning("spmm_brick_diag", mode)
    run(f"  cp.async rings, {name}")
ctx.set_tuning("spmm_brick_diag", 0)
if len(sys.argv) > 2 and sys.argv[2] == "bricks":
    sys.exit(0)
ctx.set_tuning("spmm_brick", 0)
run("plain, tile-major (8,8), block 256 (round-2 default)")
for tile in ((2, 2), (4, 4), (2, 4), (4, 2)):
    order, hints = fem.brick_schedule(x, n, msh._lattice, tile=tile)
    per_box = int(np.diff(hints).max())
    for block in (256, 512, 1024):
        ctx.set_tuning("spmm_block", block)
        ctx.set_slice_order(L.PAT_VV, order)
        run(f"plain, box schedule tile {tile} ({per_box} slices/box), unpadded, block {block}")
        for group in sorted({block // 32, 32}):
            ctx.set_slice_order(L.PAT_VV, fem.pad_order(order, hints, group))
            run(f"plain, box schedule tile {tile}, padded to {group}, block {block}")
