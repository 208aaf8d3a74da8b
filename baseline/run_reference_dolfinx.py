#!/usr/bin/env python3
"""Time the REAL reference (oasisx on DOLFINx/PETSc) on bench.py's workload -- for a machine that has the FEniCSx stack.

    mpirun -n <cores> python baseline/run_reference_dolfinx.py [--mesh 96] [--steps 100] [--warmup 3]

NOT RUN in this repository's environment: neither the build container nor the GPU boxes can install
fenics-dolfinx / petsc4py / mpi4py, which is why ``bench.py --impl reference`` times the C++/OpenMP
restatement instead.  This script exists so that anyone with the stack can put the true number next to ours: it
builds the z-extruded 3D Taylor-Green problem of SURVEY.md 8(d) with the reference's own public API
(``oasisx.FractionalStep_AB_CN``, ``oasisx.DirichletBC``; set-up as in ``demo/taylor_green.py:126-182`` of the
reference, extended to 3D) and the Krylov options of bench.py, times ``solver.solve(dt, nu, max_iter=1)``
(``fracstep.py:660``) with ``time.perf_counter`` around the loop, max over ranks, and prints one JSON line in bench.py's
format with ``"impl": "reference-dolfinx"``.

Notes for whoever runs it:
  * for ``bcs_p=[]`` the reference forces a MUMPS direct solve of the pressure system (``fracstep.py:562-576``)
    whatever ``solver_options["pressure"]`` says; at 96^3 (0.9 M pressure dofs) that factorisation is feasible;
  * ``b200_*`` keys are options of this repository's solvers; PETSc ignores them, so they are not passed;
  * ``low_memory_version=False`` as in the reference demo (``demo/taylor_green.py:115``).
"""
from __future__ import annotations

import argparse
import json
import os
import time

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, default=96)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dt", type=float, default=0.005)
    ap.add_argument("--nu", type=float, default=0.01)
    args = ap.parse_args()

    import dolfinx
    import oasisx
    from mpi4py import MPI

    comm = MPI.COMM_WORLD
    N, dt, nu = args.mesh, args.dt, args.nu
    mesh = dolfinx.mesh.create_box(comm, [[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]], [N, N, N],
                                   cell_type=dolfinx.mesh.CellType.tetrahedron)
    fdim = mesh.topology.dim - 1
    mesh.topology.create_connectivity(fdim, fdim + 1)
    facets = np.sort(dolfinx.mesh.exterior_facet_indices(mesh.topology))
    tag = np.int32(3)
    facet_tags = dolfinx.mesh.meshtags(mesh, fdim, facets, np.full_like(facets, tag, dtype=np.int32))

    class Field:
        """z-extruded Taylor-Green vortex: (u, v, 0)(x, y, t), p(x, y, t)."""

        t_u = 0.0
        t_p = 0.0

        def ux(self, x):
            return -np.cos(np.pi * x[0]) * np.sin(np.pi * x[1]) * np.exp(-2.0 * nu * np.pi**2 * self.t_u)

        def uy(self, x):
            return np.cos(np.pi * x[1]) * np.sin(np.pi * x[0]) * np.exp(-2.0 * nu * np.pi**2 * self.t_u)

        def uz(self, x):
            return np.zeros_like(x[0])

        def p(self, x):
            return -0.25 * (np.cos(2 * np.pi * x[0]) + np.cos(2 * np.pi * x[1])) * np.exp(-4 * nu * np.pi**2 * self.t_p)

    fld = Field()
    comps = [fld.ux, fld.uy, fld.uz]
    bcs_u = [[oasisx.DirichletBC(f, oasisx.LocatorMethod.TOPOLOGICAL, (facet_tags, tag))] for f in comps]
    krylov = {
        "tentative": {"ksp_type": "bcgs", "pc_type": "jacobi", "ksp_rtol": 1e-10, "ksp_initial_guess_nonzero": True},
        "pressure": {"ksp_type": "cg", "pc_type": "gamg", "ksp_rtol": 1e-10},  # overridden by MUMPS when bcs_p == [] (see above)
        "scalar": {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-10, "ksp_initial_guess_nonzero": True},
    }
    t_setup = time.perf_counter()
    solver = oasisx.FractionalStep_AB_CN(mesh, ("Lagrange", 2), ("Lagrange", 1), bcs_u=bcs_u, bcs_p=[], rotational=False,
                                         solver_options=krylov, options={"low_memory_version": False}, body_force=None)
    fld.t_u = -dt
    for i, f in enumerate(comps):
        solver._u2[i].interpolate(f)
    fld.t_u = 0.0
    for i, f in enumerate(comps):
        solver._u1[i].interpolate(f)
    fld.t_p = -dt / 2
    solver._p.interpolate(fld.p)
    t_setup = time.perf_counter() - t_setup

    def step():
        fld.t_u += dt
        fld.t_p += dt
        solver.solve(dt, nu, max_iter=1)

    for _ in range(max(args.warmup, 0)):
        step()
    comm.Barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    sec = comm.allreduce((time.perf_counter() - t0) / args.steps, op=MPI.MAX)
    n_dofs = 3 * solver._Vi[0][0].dofmap.index_map.size_global + solver._Q.dofmap.index_map.size_global
    if comm.rank == 0:
        print(json.dumps({
            "impl": "reference-dolfinx", "metric": "IPCS steps/s, 3D Taylor-Green P2-P1 box", "value": 1.0 / sec, "unit": "steps/s",
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True, "dtype": "f64",
            "config": {"workload": f"3D Taylor-Green P2-P1 {N}^3 box (z-extruded exact solution), dt={dt}, nu={nu}, max_iter=1",
                       "mesh": N, "dofs": int(n_dofs), "krylov": krylov, "setup_s": t_setup},
            "cpu_baseline": {"kind": "reference", "cores": comm.size, "host_cpus": os.cpu_count(),
                             "dolfinx": dolfinx.__version__},
        }))


if __name__ == "__main__":
    main()
