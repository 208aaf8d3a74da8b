"""Shared problem set-ups for the parity tests: the Taylor-Green vortex of
``/root/reference/demo/taylor_green.py:36-53,126-182`` (2D, and its z-extruded 3D version,
SURVEY.md F5) built both for the CUDA path (``oasisx_b200``) and for the oracle."""
from __future__ import annotations

import numpy as np

from oasisx_b200 import fem, mesh as bmesh


class TaylorGreen:
    """The exact solution is separable, f(x) * g(t).  A caller that passes the SAME coordinate array again (the
    boundary conditions do, every time step) gets the spatial factor from a one-entry cache and only the scalar time
    factor is re-evaluated -- bitwise the same values as evaluating the full expression (same operation order), the
    way the reference's compiled ``Expression`` with a time ``Constant`` avoids re-deriving anything per step."""

    def __init__(self, nu: float, gdim: int):
        self.nu, self.gdim = nu, gdim
        self.t_u = 0.0
        self.t_p = 0.0
        self._cache = {}

    def _spatial(self, key, x, f):
        hit = self._cache.get(key)
        if hit is None or hit[0] is not x:
            hit = (x, f(x))
            self._cache[key] = hit
        return hit[1]

    def eval_x(self, x):
        return self._spatial("x", x, lambda x: -np.cos(np.pi * x[0]) * np.sin(np.pi * x[1])) * np.exp(-2.0 * self.nu * np.pi**2 * self.t_u)

    def eval_y(self, x):
        return self._spatial("y", x, lambda x: np.cos(np.pi * x[1]) * np.sin(np.pi * x[0])) * np.exp(-2.0 * self.nu * np.pi**2 * self.t_u)

    def eval_z(self, x):
        return np.zeros_like(x[0])

    def eval_p(self, x):
        return self._spatial("p", x, lambda x: -0.25 * (np.cos(2 * np.pi * x[0]) + np.cos(2 * np.pi * x[1]))) * np.exp(-4 * self.nu * np.pi**2 * self.t_p)

    @property
    def components(self):
        return [self.eval_x, self.eval_y, self.eval_z][: self.gdim]


def make_mesh(gdim: int, N: int, comm=None):
    if gdim == 2:
        return bmesh.create_rectangle(comm, [[-1.0, -1.0], [1.0, 1.0]], [N, N])
    return bmesh.create_box(comm, [[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]], [N, N, N])


def boundary_facets(msh):
    d = msh.topology.dim
    msh.topology.create_connectivity(d - 1, d)
    return bmesh.exterior_facet_indices(msh.topology)


def make_oracle(msh, deg_u, tg: TaylorGreen, dt, **kw):
    from oracle.ipcs_oracle import OracleIPCS

    d = msh.geometry.dim
    V = fem.functionspace(msh, ("Lagrange", deg_u))
    Q = fem.functionspace(msh, ("Lagrange", 1))
    bd = fem.locate_dofs_topological(V, d - 1, boundary_facets(msh))
    o = OracleIPCS(msh.geometry.x, msh.geometry.dofmap, d, V.dofmap.list, Q.dofmap.list,
                   V.tabulate_dof_coordinates(), Q.tabulate_dof_coordinates(), deg_u,
                   bcs_u=[[(bd, f)] for f in tg.components], **kw)
    xV, xQ = V.tabulate_dof_coordinates().T, Q.tabulate_dof_coordinates().T
    tg.t_u = -dt
    o.u2 = [f(xV) for f in tg.components]
    tg.t_u = 0.0
    o.u1 = [f(xV) for f in tg.components]
    tg.t_p = -dt / 2
    o.p = tg.eval_p(xQ)
    return o


def make_solver(msh, deg_u, tg: TaylorGreen, dt, solver_options=None, low_memory=False, **kw):
    """The set-up of ``demo/taylor_green.py:135-182`` against oasisx_b200."""
    import oasisx_b200 as oasisx

    d = msh.geometry.dim
    facets = boundary_facets(msh)
    value = np.int32(3)
    tags = bmesh.meshtags(msh, d - 1, facets, np.full_like(facets, value, dtype=np.int32))
    bcs_u = [[oasisx.DirichletBC(f, oasisx.LocatorMethod.TOPOLOGICAL, (tags, value))] for f in tg.components]
    if solver_options is None:
        lu = {"ksp_type": "preonly", "pc_type": "lu"}
        solver_options = {"tentative": lu, "pressure": lu, "scalar": lu}
    s = oasisx.FractionalStep_AB_CN(msh, ("Lagrange", deg_u), ("Lagrange", 1), bcs_u=bcs_u, bcs_p=[],
                                    solver_options=solver_options, options={"low_memory_version": low_memory}, **kw)
    tg.t_u = -dt
    for i, f in enumerate(tg.components):
        s._u2[i].interpolate(f)
    tg.t_u = 0.0
    for i, f in enumerate(tg.components):
        s._u1[i].interpolate(f)
    tg.t_p = -dt / 2
    s._p.interpolate(tg.eval_p)
    return s


def relerr(a, b, scale=None):
    """max|a-b| / scale, scale = max|b| unless given (pass the max over all components when one
    component is identically zero, e.g. w in the z-extruded Taylor-Green field)."""
    a, b = np.asarray(a), np.asarray(b)
    if scale is None:
        scale = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / max(scale, 1e-300))


def vscale(vs):
    return max(float(np.max(np.abs(v))) for v in vs)


def make_cpu_port(msh, deg_u, tg: TaylorGreen, dt, **kw):
    """The same set-up for the C++/OpenMP CPU restatement (oracle/ipcs_cpu.cpp), which reaches meshes the LU oracle
    cannot and is itself pinned against that oracle in tests/test_cpu_port.py."""
    from oracle import ipcs_cpu as cpu

    d = msh.geometry.dim
    V, Q = fem.functionspace(msh, ("Lagrange", deg_u)), fem.functionspace(msh, ("Lagrange", 1))
    bd = fem.locate_dofs_topological(V, d - 1, boundary_facets(msh))
    c = cpu.CpuIPCS(msh.geometry.x, msh.geometry.dofmap, d, V.dofmap.list, Q.dofmap.list, V.tabulate_dof_coordinates(),
                    Q.tabulate_dof_coordinates(), deg_u, bcs_u=[[(bd, f)] for f in tg.components], **kw)
    xV, xQ = V.tabulate_dof_coordinates().T, Q.tabulate_dof_coordinates().T
    tg.t_u = -dt
    for i, f in enumerate(tg.components):
        c.set(cpu.U2, i, f(xV))
    tg.t_u = 0.0
    for i, f in enumerate(tg.components):
        c.set(cpu.U1, i, f(xV))
    tg.t_p = -dt / 2
    c.set(cpu.P, 0, tg.eval_p(xQ))
    return c
