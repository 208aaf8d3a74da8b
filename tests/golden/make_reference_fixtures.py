#!/usr/bin/env python3
"""Golden-fixture generator: run the REAL reference (oasisx on DOLFINx/PETSc) and dump what pins the parity.

    python tests/golden/make_reference_fixtures.py            (serial; needs fenics-dolfinx >= 0.10, petsc4py, oasisx)

NOT RUN in this repository's environment (no FEniCSx stack: DESIGN.md section 2) -- which is why the oracle is
"parity unpinned".  Whoever has the stack runs this once and commits the two ``.npz`` files it writes next to this
script; ``tests/test_golden_reference.py`` then compares the numpy oracle AND the CUDA path with them (dofs are
matched by their coordinates, so DOLFINx's dof numbering does not matter).  Content, for the 2D Taylor-Green problem
of ``demo/taylor_green.py`` on 8 x 8 and the z-extruded 3D one on 4 x 4 x 4 (P2-P1, dt = 0.005, nu = 0.01, LU solves):

    xV, xQ          dof coordinates of the velocity-component space and of the pressure space
    A               dense copy of ``solver._A`` after ``assemble_first`` of the first step (``fracstep.py:411-472``)
    b_first, rhs1   per component after ``assemble_first`` / ``velocity_tentative_assemble`` (``:449-506``)
    u_k, p_k        fields after each of 3 calls of ``solver.solve(dt, nu, max_iter=1)`` (``:660-696``)
    err_u, err_p    the demo's L2 error functionals after each step (``demo/taylor_green.py:186-207``)
"""
from __future__ import annotations

import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DT, NU, STEPS = 0.005, 0.01, 3


class U:
    def __init__(self, nu):
        self.nu, self.t = nu, 0.0

    def eval_x(self, x):
        return -np.cos(np.pi * x[0]) * np.sin(np.pi * x[1]) * np.exp(-2.0 * self.nu * np.pi**2 * self.t)

    def eval_y(self, x):
        return np.cos(np.pi * x[1]) * np.sin(np.pi * x[0]) * np.exp(-2.0 * self.nu * np.pi**2 * self.t)

    def eval_z(self, x):
        return np.zeros_like(x[0])


def p_exact(x, t, nu):
    return -0.25 * (np.cos(2 * np.pi * x[0]) + np.cos(2 * np.pi * x[1])) * np.exp(-4 * nu * np.pi**2 * t)


def run(gdim: int, N: int, path: str):
    from mpi4py import MPI

    import dolfinx
    import oasisx
    import ufl

    if gdim == 2:
        mesh = dolfinx.mesh.create_rectangle(MPI.COMM_WORLD, [np.array([-1.0, -1.0]), np.array([1.0, 1.0])], [N, N],
                                             dolfinx.mesh.CellType.triangle)
    else:
        mesh = dolfinx.mesh.create_box(MPI.COMM_WORLD, [np.array([-1.0] * 3), np.array([1.0] * 3)], [N, N, N],
                                       dolfinx.mesh.CellType.tetrahedron)
    assert mesh.comm.size == 1, "generate the fixtures in serial"
    fdim = mesh.topology.dim - 1
    mesh.topology.create_connectivity(fdim, fdim + 1)
    facets = dolfinx.mesh.exterior_facet_indices(mesh.topology)
    value = np.int32(3)
    tags = dolfinx.mesh.meshtags(mesh, fdim, np.sort(facets), np.full_like(facets, value, dtype=np.int32))
    u_ex = U(NU)
    comps = [u_ex.eval_x, u_ex.eval_y, u_ex.eval_z][:gdim]
    bcs_u = [[oasisx.DirichletBC(f, oasisx.LocatorMethod.TOPOLOGICAL, (tags, value))] for f in comps]
    lu = {"ksp_type": "preonly", "pc_type": "lu", "pc_factor_mat_solver_type": "mumps"}
    solver = oasisx.FractionalStep_AB_CN(mesh, ("Lagrange", 2), ("Lagrange", 1), bcs_u=bcs_u, bcs_p=[],
                                         solver_options={"tentative": lu, "pressure": lu, "scalar": lu},
                                         options={"low_memory_version": False})
    u_ex.t = -DT
    for i, f in enumerate(comps):
        solver._u2[i].interpolate(f)
    u_ex.t = 0.0
    for i, f in enumerate(comps):
        solver._u1[i].interpolate(f)
    solver._p.interpolate(lambda x: p_exact(x, -DT / 2, NU))
    Vi = solver._u1[0].function_space
    out = {"gdim": gdim, "N": N, "dt": DT, "nu": NU,
           "xV": Vi.tabulate_dof_coordinates()[: Vi.dofmap.index_map.size_local],
           "xQ": solver._Q.tabulate_dof_coordinates()[: solver._Q.dofmap.index_map.size_local]}
    # first step, stage by stage (test/test_tentative_velocity.py:172-174)
    u_ex.t = DT
    solver._ps.x.array[:] = solver._p.x.array[:]
    for bcl in bcs_u:
        for bc in bcl:
            bc.update_bc()
    solver.assemble_first(DT, NU)
    ip, ix, vals = solver._A.getValuesCSR()
    n = len(ip) - 1
    A = np.zeros((n, n))
    for r in range(n):
        A[r, ix[ip[r]:ip[r + 1]]] = vals[ip[r]:ip[r + 1]]
    out["A"] = A
    solver.velocity_tentative_assemble()
    for i in range(gdim):
        out[f"b_first_{i}"] = solver._b_first[i].x.array.copy()
        out[f"rhs1_{i}"] = solver._rhs1[i].x.array.copy()
    # whole steps from the same initial state (assemble_first above changed nothing that solve() does not redo)
    x = ufl.SpatialCoordinate(mesh)
    for k in range(STEPS):
        u_ex.t = (k + 1) * DT
        solver.solve(DT, NU, max_iter=1)
        for i in range(gdim):
            out[f"u{i}_{k}"] = solver._u[i].x.array.copy()
        out[f"p_{k}"] = solver._p.x.array.copy()
        t, tp = (k + 1) * DT, (k + 0.5) * DT
        ue = ufl.as_vector([-ufl.cos(ufl.pi * x[0]) * ufl.sin(ufl.pi * x[1]) * np.exp(-2 * NU * np.pi**2 * t),
                            ufl.cos(ufl.pi * x[1]) * ufl.sin(ufl.pi * x[0]) * np.exp(-2 * NU * np.pi**2 * t)] + ([0.0] if gdim == 3 else []))
        pe = -0.25 * (ufl.cos(2 * ufl.pi * x[0]) + ufl.cos(2 * ufl.pi * x[1])) * np.exp(-4 * NU * np.pi**2 * tp)
        du = solver.u - ue
        out[f"err_u_{k}"] = dolfinx.fem.assemble_scalar(dolfinx.fem.form(ufl.inner(du, du) * ufl.dx))
        out[f"err_p_{k}"] = dolfinx.fem.assemble_scalar(dolfinx.fem.form((solver._p - pe) ** 2 * ufl.dx))
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    run(2, 8, os.path.join(HERE, "reference_tg2d_8.npz"))
    run(3, 4, os.path.join(HERE, "reference_tg3d_4.npz"))
