#!/usr/bin/env python3
# This driver follows demo/taylor_green.py of Oasisx (argument list, exact-solution classes, loop and rate tail):
#   Copyright (C) 2022 Jørgen Schartum Dokken -- This file is part of Oasisx -- SPDX-License-Identifier: MIT
# with every DOLFINx / UFL / PETSc call replaced by oasisx_b200.
"""Taylor-Green convergence demo on B200 -- the driver of /root/reference/demo/taylor_green.py with
`oasisx` replaced by `oasisx_b200` and the DOLFINx mesh/tag calls by the built-in provider.  Same
command line (-N, -T0, -T1, -dt, -nu, -u, -p, -lm, -r), same initial/boundary data, same error
functionals and convergence rates; plus -d 3 for the z-extruded 3D field.

    python demo/taylor_green.py -N 8 -N 16 -N 32 -dt 0.005        (the reference CI invocation)
"""
import argparse
import logging
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oasisx_b200 as oasisx  # noqa: E402
from oasisx_b200 import mesh as bmesh  # noqa: E402


class U:
    def __init__(self, t, nu):
        self.t, self.nu = t, nu

    def eval_x(self, x):
        return -np.cos(np.pi * x[0]) * np.sin(np.pi * x[1]) * np.exp(-2.0 * self.nu * np.pi**2 * float(self.t))

    def eval_y(self, x):
        return np.cos(np.pi * x[1]) * np.sin(np.pi * x[0]) * np.exp(-2.0 * self.nu * np.pi**2 * float(self.t))

    def eval_z(self, x):
        return np.zeros_like(x[0])

    def trig_terms(self, which):
        """The same field as trigonometric product terms: evaluated on the device by assemble_l2_error_sq."""
        g, pi = np.exp(-2.0 * self.nu * np.pi**2 * float(self.t)), np.pi
        return [[-g, pi, 0, 0, 0, 0, pi, 0, 0, 2, 1, 0], [g, pi, 0, 0, 0, 0, pi, 0, 0, 1, 2, 1]]


class Pexact:
    def __init__(self, t, nu):
        self.t, self.nu = t, nu

    def __call__(self, x):
        return -0.25 * (np.cos(2 * np.pi * x[0]) + np.cos(2 * np.pi * x[1])) * np.exp(-4 * np.pi**2 * self.nu * float(self.t))

    def trig_terms(self, which):
        g, pi = -0.25 * np.exp(-4 * np.pi**2 * self.nu * float(self.t)), np.pi
        return [[g, 2 * pi, 0, 0, 0, 0, 0, 0, 0, 2, 0, 0], [g, 0, 2 * pi, 0, 0, 0, 0, 0, 0, 2, 0, 0]]


parser = argparse.ArgumentParser(description="Taylor-Green convergence demo", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
parser.add_argument("-N", "--refinement", type=int, dest="Ns", action="append", required=True,
                    help="The number of elements in x and y direction")
parser.add_argument("-T0", "--T-start", dest="T_start", type=float, default=0, help="Start time of simulation")
parser.add_argument("-T1", "--T-end", dest="T_end", type=float, default=1, help="End time of simulation")
parser.add_argument("-dt", dest="dt", type=float, default=0.1, help="Time step")
parser.add_argument("-nu", dest="nu", type=float, default=0.01, help="Kinematic viscosity")
parser.add_argument("-u", dest="u_deg", type=int, default=2, help="Degree of velocity space")
parser.add_argument("-p", dest="p_deg", type=int, default=1, help="Degree of pressure space")
parser.add_argument("-lm", "--low-memory", dest="lm", action="store_true", default=False)
parser.add_argument("-r", "--rotational", dest="rot", action="store_true", default=False)
parser.add_argument("-d", "--dim", dest="dim", type=int, default=2, choices=[2, 3], help="2: the reference demo; 3: z-extruded field")
parser.add_argument("--krylov", action="store_true", help="Krylov solvers (BiCGStab / CG + multigrid / CG, rtol 1e-10) instead of the "
                                                          "'preonly + lu' of the reference: what large 3D meshes need")
parser.add_argument("--host-errors", action="store_true", help="sample the exact fields on the host for the error functionals "
                                                               "(the reference's way; moves cells x points x components doubles per step)")
parser.add_argument("--export", default=None, help="write <prefix>_N<N>.vtu with u and p after the last step (VTXWriter stand-in)")
parser.add_argument("--checkpoint", default=None, help="write <prefix>_N<N>.npz with u1, u2, p and t after the last step")
inputs = parser.parse_args()
logger = logging.getLogger("Oasisx")

dt, nu = inputs.dt, inputs.nu
assert inputs.T_start < inputs.T_end
T_end, T_start = inputs.T_end, inputs.T_start
num_steps = int((T_end - T_start) // dt)
assert inputs.u_deg > inputs.p_deg
el_u, el_p = ("Lagrange", inputs.u_deg), ("Lagrange", inputs.p_deg)
options = {"low_memory_version": inputs.lm}
solver_options = {
    "tentative": {"ksp_type": "preonly", "pc_type": "lu"},
    "pressure": {"ksp_type": "preonly", "pc_type": "lu"},
    "scalar": {"ksp_type": "preonly", "pc_type": "lu"},
}
if inputs.krylov:
    kry = {"ksp_rtol": 1e-10, "ksp_initial_guess_nonzero": True}
    solver_options = {"tentative": {"ksp_type": "bcgs", "pc_type": "jacobi", "b200_guess": "extrapolate2", **kry},
                      "pressure": {"ksp_type": "cg", "pc_type": "mg", "b200_guess": "extrapolate", **kry},
                      "scalar": {"ksp_type": "cg", "pc_type": "jacobi", "b200_guess": "extrapolate2", **kry}}
gdim = inputs.dim
space_errors = np.zeros((2, len(inputs.Ns)))
hs = np.zeros(len(inputs.Ns))
for n, N in enumerate(inputs.Ns):
    if gdim == 2:
        mesh = bmesh.create_rectangle(None, [[-1, -1], [1, 1]], [N, N])
    else:
        mesh = bmesh.create_box(None, [[-1, -1, -1], [1, 1, 1]], [N, N, N])
    dim = mesh.topology.dim - 1
    mesh.topology.create_connectivity(dim, dim + 1)
    facets = bmesh.exterior_facet_indices(mesh.topology)
    value = np.int32(3)
    values = np.full_like(facets, value, dtype=np.int32)
    sort = np.argsort(facets)
    facet_tags = bmesh.meshtags(mesh, dim, facets[sort], values[sort])

    class Time:  # stand-in for dolfinx.fem.Constant holding the current time
        def __init__(self, v):
            self.value = v

        def __float__(self):
            return float(self.value)

    u_time, p_time = Time(T_start), Time(T_start - dt / 2.0)
    u_ex = U(t=u_time, nu=nu)
    p_ex = Pexact(p_time, nu)
    comps = [u_ex.eval_x, u_ex.eval_y, u_ex.eval_z][:gdim]
    bcs_u = [[oasisx.DirichletBC(f, oasisx.LocatorMethod.TOPOLOGICAL, (facet_tags, value))] for f in comps]
    solver = oasisx.FractionalStep_AB_CN(mesh, el_u, el_p, bcs_u=bcs_u, bcs_p=[], rotational=inputs.rot,
                                         solver_options=solver_options, options=options, body_force=None)
    u_time.value = T_start - dt
    for i, f in enumerate(comps):
        solver._u2[i].interpolate(f)
    u_time.value = T_start
    for i, f in enumerate(comps):
        solver._u1[i].interpolate(f)
    solver._p.interpolate(p_ex)

    error_space_time = np.zeros((2, num_steps))
    u_time.value = T_start
    for i in range(num_steps):
        u_time.value += dt
        p_time.value += dt
        solver.solve(dt, nu, max_iter=1)
        # demo/taylor_green.py:204-207; the analytic fields are evaluated on the device unless --host-errors
        error_u = mesh.comm.allreduce(solver.assemble_l2_error_sq("u", comps if inputs.host_errors else u_ex, degree=10))
        error_p = mesh.comm.allreduce(solver.assemble_l2_error_sq("p", p_ex.__call__ if inputs.host_errors else p_ex, degree=10))
        error_space_time[:, i] = [error_u, error_p]
    if inputs.export:
        from oasisx_b200.io import write_vtu

        Vs = solver._Vi[0][0]
        write_vtu(f"{inputs.export}_N{N}.vtu", Vs, {"u": solver.u.x.array.reshape(-1, gdim)})
        write_vtu(f"{inputs.export}_N{N}_p.vtu", solver._Q, {"p": solver._p.x.array})
    if inputs.checkpoint:
        from oasisx_b200.io import save_checkpoint

        save_checkpoint(f"{inputs.checkpoint}_N{N}", solver, float(u_time.value))
    hmax = mesh.comm.allreduce(np.max(mesh.h(mesh.topology.dim, np.arange(mesh.num_cells, dtype=np.int32))), op="max")
    space_time_u_L2 = np.sqrt(dt * np.sum(error_space_time[0, :]))
    space_time_p_L2 = np.sqrt(dt * np.sum(error_space_time[1, :]))
    print(f"N={N} hmax={hmax:.6f} space_time_u_L2={space_time_u_L2:.6e} space_time_p_L2={space_time_p_L2:.6e}", flush=True)
    hs[n] = hmax
    space_errors[:, n] = [space_time_u_L2, space_time_p_L2]

order = np.argsort(hs)[::-1]
hs[:] = hs[order]
space_errors[0, :] = space_errors[0, order]
space_errors[1, :] = space_errors[1, order]
if len(hs) > 1:
    rate_u = np.log(space_errors[0, 1:] / space_errors[0, :-1]) / np.log(hs[1:] / hs[:-1])
    rate_p = np.log(space_errors[1, 1:] / space_errors[1, :-1]) / np.log(hs[1:] / hs[:-1])
    print(f"Convergence rates u: {rate_u}")
    print(f"Convergence rates p: {rate_p}")
