"""Mirror of /root/reference/test/test_tentative_velocity.py:87-242: unit square 10x10, inlet callable
BC + walls + outlet PressureBC(4.0), +/- body force; `assemble_first`, `velocity_tentative_assemble`,
`velocity_tentative_solve` through oasisx_b200, compared with the direct statement of the equation (the
CPU oracle) -- RHS vectors (the reference's assertion, :235), the matrix (the reference's dead check,
:227-229, made live) and the solution; then whole time steps incl. the pressure Dirichlet rows."""
import numpy as np
import pytest
import scipy.sparse as sp

from oasisx_b200 import DirichletBC, FractionalStep_AB_CN, LocatorMethod, PressureBC, fem, mesh as bmesh
from problems import relerr, vscale

pytestmark = pytest.mark.gpu


class Inlet:
    def __init__(self, t):
        self.t = t

    def eval(self, x):
        return (1 + self.t) * np.sin(np.pi * x[1])


def build(deg_u, body_force, solver_options=None, low_memory=False, comm=None, device=0, foreign=None):
    from oracle.ipcs_oracle import OracleIPCS

    def tagged(m):
        dim = m.topology.dim - 1
        left = bmesh.locate_entities_boundary(m, dim, lambda x: np.isclose(x[0], 0))
        tb = bmesh.locate_entities_boundary(m, dim, lambda x: np.logical_or(np.isclose(x[1], 0), np.isclose(x[1], 1)))
        right = bmesh.locate_entities_boundary(m, dim, lambda x: np.isclose(x[0], 1))
        facets = np.hstack([left, tb, right])
        values = np.hstack([np.full_like(left, 1), np.full_like(tb, 2), np.full_like(right, 3)])
        order = np.argsort(facets)
        return dim, left, tb, right, bmesh.meshtags(m, dim, facets[order], values[order])

    # several ranks: the solver gets this rank's slab of the mesh (oasisx_b200.slab), the single-process oracle the
    # whole mesh
    smsh = bmesh.create_unit_square(comm, 10, 10)
    msh = smsh if comm is None else bmesh.create_unit_square(None, 10, 10)
    tags = tagged(smsh)[4]
    dim, left, tb, right, _ = tagged(msh)
    if foreign is not None:
        # the solver sees the mesh as a DOLFINx mesh would arrive (tests/fake_dolfinx.py): oasisx_b200.adapter route,
        # facets from the mesh's cell-to-facet connectivity, dofs located by "DOLFINx"
        import sys

        from fake_dolfinx import make_fake

        mod, fmesh, _ = make_fake(msh, deg_u, 1, 1, 0)
        foreign.setitem(sys.modules, "dolfinx", mod)
        smsh, tags = fmesh, fmesh.fake_tags(tags)
    inlet = Inlet(0)
    bc_tb = DirichletBC(0.0, LocatorMethod.TOPOLOGICAL, (tags, 2))
    bc_inlet_x = DirichletBC(inlet.eval, LocatorMethod.TOPOLOGICAL, (tags, 1))
    bc_inlet_y = DirichletBC(0.0, LocatorMethod.TOPOLOGICAL, (tags, 1))
    bcs_u = [[bc_inlet_x, bc_tb], [bc_inlet_y, bc_tb]]
    bcs_p = [PressureBC(4.0, (tags, 3))]
    f = np.array([0.3, -0.1]) if body_force else None
    lu = {"ksp_type": "preonly", "pc_type": "lu"}
    s = FractionalStep_AB_CN(smsh, ("Lagrange", deg_u), ("Lagrange", 1), bcs_u=bcs_u, bcs_p=bcs_p,
                             solver_options=solver_options or {"tentative": lu, "pressure": lu, "scalar": lu},
                             options={"low_memory_version": low_memory}, body_force=f, device=device)
    V, Q = fem.functionspace(msh, ("Lagrange", deg_u)), fem.functionspace(msh, ("Lagrange", 1))
    dl, dtb = fem.locate_dofs_topological(V, dim, left), fem.locate_dofs_topological(V, dim, tb)
    pdofs = fem.locate_dofs_topological(Q, dim, right)
    o = OracleIPCS(msh.geometry.x, msh.geometry.dofmap, 2, V.dofmap.list, Q.dofmap.list, V.tabulate_dof_coordinates(),
                   Q.tabulate_dof_coordinates(), deg_u,
                   bcs_u=[[(dl, inlet.eval), (dtb, 0.0)], [(dl, 0.0), (dtb, 0.0)]], bcs_p=[pdofs], body_force=f,
                   pressure_facets=[_global_pressure_facets(msh, right) + (4.0,)])
    return s, o, inlet, bc_inlet_x


def _global_pressure_facets(msh, facets):
    """(cell, local facet index) of the tagged facets in GLOBAL cell numbering (what the single-process oracle needs;
    a multi-rank PressureBC holds its slab's share in local numbering)."""
    fdim = msh.topology.dim - 1
    msh.topology.create_connectivity(fdim, msh.topology.dim)
    cf = msh.topology.cell_entities(fdim)
    cells, local = np.nonzero(np.isin(cf, facets))
    return cells.astype(np.int32), local.astype(np.int32)


def test_channel_with_pressure_bc_on_a_foreign_mesh(monkeypatch):
    """The open channel (two Dirichlet BCs per component + PressureBC) with the mesh consumed through the DOLFINx
    adapter: two inner-iterated steps against the oracle."""
    kry = {"ksp_type": "bcgs", "pc_type": "jacobi", "ksp_rtol": 1e-12}
    cg = {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-12}
    s, o, inlet, _ = build(2, True, solver_options={"tentative": kry, "pressure": cg, "scalar": cg}, foreign=monkeypatch)
    assert s._foreign and len(s._bcs_p[0]._facet_cells) == 10
    dt, nu = 0.01, 0.5
    inlet.t = 0.0
    for n in range(2):
        inlet.t += dt
        s.solve(dt, nu, max_iter=2, max_error=1e-30)
        o.solve(dt, nu, max_iter=2, max_error=1e-30)
        for i in range(2):
            assert relerr(s._u[i].x.array_ro(), o.u[i], vscale(o.u)) <= 1e-7
        assert relerr(s._p.x.array_ro(), o.p) <= 1e-7
    assert max(np.abs(p).max() for p in o.p_surf) > 1e-3  # the natural pressure term is there and matters


@pytest.mark.parametrize("body_force", [True, False])
@pytest.mark.parametrize("low_memory", [True, False])
@pytest.mark.parametrize("deg_u", [1, 2])
def test_tentative(deg_u, low_memory, body_force):
    s, o, inlet, bc_inlet_x = build(deg_u, body_force, low_memory=low_memory)
    dt, nu = 0.1, 0.5
    xV = o.xV.T
    inlet.t = -2 * dt
    for i in range(2):
        s._u2[i].interpolate(inlet.eval)
        o.u2[i] = inlet.eval(xV)
    inlet.t = -dt
    for i in range(2):
        s._u1[i].interpolate(inlet.eval)
        o.u1[i] = inlet.eval(xV)
    inlet.t = dt
    bc_inlet_x.update_bc()
    o.update_bcs()
    s._ps.interpolate(lambda x: x[1])
    o.ps = o.xQ[:, 1].copy()
    s.assemble_first(dt, nu)
    s.velocity_tentative_assemble()
    diff, reasons = s.velocity_tentative_solve()
    o.assemble_first(dt, nu)
    o.velocity_tentative_assemble()
    o.velocity_tentative_solve()
    # the natural pressure term is there and matters
    assert max(np.abs(p).max() for p in o.p_surf) > 1e-3
    ip, ix, v = s._A.getValuesCSR()
    A = sp.csr_matrix((v, ix, ip), shape=s._A.getSize())
    assert abs(A - o.A).max() <= 1e-12 * abs(o.A).max()  # test_tentative_velocity.py:227-229
    assert (reasons > 0).all()  # :234
    for i in range(2):
        assert relerr(s._rhs1[i].x.array_ro(), o.rhs1[i], vscale(o.rhs1)) <= 1e-12  # :235
        assert relerr(s._u[i].x.array_ro(), o.u[i], vscale(o.u)) <= 1e-9
    # the stages the reference leaves commented out (:237-240)
    s.pressure_assemble(dt)
    assert s.pressure_solve(nu) > 0
    assert (s.velocity_update(dt) > 0).all()
    o.pressure_assemble(dt)
    o.pressure_solve(nu)
    o.velocity_update(dt)
    assert relerr(s._dp.x.array_ro(), o.dp) <= 1e-8
    for i in range(2):
        assert relerr(s._u[i].x.array_ro(), o.u[i], vscale(o.u)) <= 1e-8


def test_channel_time_steps_with_pressure_bc():
    s, o, inlet, _ = build(2, False)
    dt, nu = 0.01, 0.5
    inlet.t = 0.0
    for n in range(3):
        inlet.t += dt
        s.solve(dt, nu, max_iter=2, max_error=1e-30)
        o.solve(dt, nu, max_iter=2, max_error=1e-30)
        for i in range(2):
            assert relerr(s._u[i].x.array_ro(), o.u[i], vscale(o.u)) <= 1e-8, n
        assert relerr(s._p.x.array_ro(), o.p) <= 1e-8, n


def test_pressure_bc_constant_is_read_live():
    """A `fem.Constant` outlet pressure changed between two `assemble_first` calls changes the assembled surface term
    (the reference keeps the Constant inside the ds-form, bcs.py:233-242; round-1 advisor finding: it was frozen)."""
    msh = bmesh.create_unit_square(None, 6, 6)
    dim = msh.topology.dim - 1
    right = bmesh.locate_entities_boundary(msh, dim, lambda x: np.isclose(x[0], 1))
    left = bmesh.locate_entities_boundary(msh, dim, lambda x: np.isclose(x[0], 0))
    facets = np.hstack([left, right])
    values = np.hstack([np.full_like(left, 1), np.full_like(right, 3)])
    order = np.argsort(facets)
    tags = bmesh.meshtags(msh, dim, facets[order], values[order])
    pc = fem.Constant(msh, 4.0)
    bcs_u = [[DirichletBC(1.0, LocatorMethod.TOPOLOGICAL, (tags, 1))], [DirichletBC(0.0, LocatorMethod.TOPOLOGICAL, (tags, 1))]]
    lu = {"ksp_type": "preonly", "pc_type": "lu"}
    s = FractionalStep_AB_CN(msh, ("Lagrange", 2), ("Lagrange", 1), bcs_u=bcs_u, bcs_p=[PressureBC(pc, (tags, 3))],
                             solver_options={"tentative": lu, "pressure": lu, "scalar": lu}, options={"low_memory_version": False})
    s.assemble_first(0.1, 0.5)
    b4 = [f.x.array_ro().copy() for f in s._b_first]
    pc.value = np.asarray(8.0)
    s.assemble_first(0.1, 0.5)
    b8 = [f.x.array_ro().copy() for f in s._b_first]
    # u1 = 0 and no body force: b_first is the surface term alone, linear in the boundary pressure
    assert max(np.abs(b).max() for b in b4) > 1e-3
    for a, b in zip(b4, b8):
        assert np.allclose(b, 2.0 * a, rtol=1e-12, atol=1e-14)


def test_dirichlet_dofs_injected_in_any_order():
    """`DirichletBC.set_dofs` with an unsorted dof array (bcs.py:103-104): values land on the right dofs (round-1
    advisor finding: the single-BC fast path assumed the sorted order of the device list)."""
    from problems import TaylorGreen, boundary_facets, make_mesh

    gdim, dt, nu = 2, 0.005, 0.01
    msh = make_mesh(gdim, 6)
    V = fem.functionspace(msh, ("Lagrange", 2))
    bd = fem.locate_dofs_topological(V, gdim - 1, boundary_facets(msh))
    tg = TaylorGreen(nu, gdim)
    lu = {"ksp_type": "preonly", "pc_type": "lu"}
    out = []
    for perm in (np.arange(len(bd)), np.random.default_rng(1).permutation(len(bd))):
        bcs_u = []
        for f in tg.components:
            bc = DirichletBC(f, LocatorMethod.GEOMETRICAL, lambda x: np.zeros(x.shape[1], bool))
            bc.set_dofs(bd[perm])
            bcs_u.append([bc])
        s = FractionalStep_AB_CN(msh, ("Lagrange", 2), ("Lagrange", 1), bcs_u=bcs_u, bcs_p=[],
                                 solver_options={"tentative": lu, "pressure": lu, "scalar": lu}, options={"low_memory_version": False})
        tg.t_u = -dt
        for i, f in enumerate(tg.components):
            s._u2[i].interpolate(f)
        tg.t_u = 0.0
        for i, f in enumerate(tg.components):
            s._u1[i].interpolate(f)
        tg.t_p = -dt / 2
        s._p.interpolate(tg.eval_p)
        tg.t_u, tg.t_p = dt, dt / 2
        s.solve(dt, nu, max_iter=1)
        out.append([f.x.array_ro().copy() for f in s._u])
    for a, b in zip(*out):
        assert np.abs(a - b).max() <= 1e-12 * max(np.abs(a).max(), 1e-30)
