"""The oracle against itself: the algebraic identities of SURVEY.md section 7 row 2 and the relational
checks the reference's own tests use (test/test_tentative_velocity.py:235,
demo/assembly_strategies.py:142)."""
import numpy as np
import pytest

from problems import TaylorGreen, make_mesh, make_oracle


@pytest.mark.parametrize("gdim,N", [(2, 6), (3, 3)])
def test_identities(gdim, N):
    tg = TaylorGreen(0.01, gdim)
    o = make_oracle(make_mesh(gdim, N), 2, tg, 0.01)
    one = np.ones(o.nV)
    vol = 2.0**gdim
    np.testing.assert_allclose(one @ (o.M @ one), vol, rtol=1e-12)
    np.testing.assert_allclose(o.K @ one, 0, atol=1e-10)
    C = o.F.convection([np.sin(o.xV[:, 0]), np.cos(o.xV[:, 1]), o.xV[:, 0] ** 2][:gdim])
    np.testing.assert_allclose(C @ one, 0, atol=1e-12)
    for i in range(gdim):
        assert abs(o.D[i] - o.P[i].T).max() < 1e-14
    np.testing.assert_allclose(o.mQ.sum(), vol, rtol=1e-12)
    np.testing.assert_allclose(o.Ap @ np.ones(o.nQ), 0, atol=1e-11)
    # polynomial exactness: P2 interpolant of a quadratic, int u^2
    u = o.xV[:, 0] ** 2 + 3 * o.xV[:, 1]
    exact = {2: 4 * (1 / 5 + 3), 3: 8 * (1 / 5 + 3)}[gdim]
    np.testing.assert_allclose(u @ (o.M @ u), exact, rtol=1e-12)


def test_matvec_rhs_equals_action_rhs():
    """demo/assembly_strategies.py:142: (M/dt - nu/2 K - C/2) u1 by matrix algebra == the same form
    integrated directly (quadrature of the action)."""
    from oracle.ipcs_oracle import simplex_quadrature, tabulate
    tg = TaylorGreen(0.3, 3)
    o = make_oracle(make_mesh(3, 3), 2, tg, 0.5)
    dt, nu = 0.5, 0.3
    u1 = np.sin(o.xV[:, 0]) * np.cos(o.xV[:, 1])
    uab = [o.xV[:, 0].copy() for _ in range(3)]
    R = o.M / dt - 0.5 * nu * o.K - 0.5 * o.F.convection(uab)
    b = R @ u1
    F = o.F
    pts, w = simplex_quadrature(3, 6)
    phi, dphi = tabulate(3, 2, pts)
    gp = np.einsum("cdk,qjd->cqjk", F.g.Kinv, dphi)
    uq = np.einsum("qa,ca->cq", phi, u1[o.vdofs])
    gu = np.einsum("cqjk,cj->cqk", gp, u1[o.vdofs])
    aq = np.stack([np.einsum("qa,ca->cq", phi, a[o.vdofs]) for a in uab], axis=2)
    integrand_v = uq / dt - 0.5 * np.einsum("cqk,cqk->cq", aq, gu)
    be = np.einsum("c,q,cq,qi->ci", F.g.detJ, w, integrand_v, phi) - 0.5 * nu * np.einsum(
        "c,q,cqk,cqik->ci", F.g.detJ, w, gu, gp)
    bd = np.zeros(o.nV)
    np.add.at(bd, o.vdofs.ravel(), be.ravel())
    np.testing.assert_allclose(b, bd, rtol=1e-11, atol=1e-12)


def test_taylor_green_2d_converges():
    """demo/taylor_green.py:225-241: space-time L2 errors fall at >= 2nd order under refinement."""
    errs = []
    for N in (8, 16):
        dt, nu = 0.005, 0.01
        tg = TaylorGreen(nu, 2)
        o = make_oracle(make_mesh(2, N), 2, tg, dt)
        eu = 0.0
        for _ in range(10):
            tg.t_u += dt
            tg.t_p += dt
            o.solve(dt, nu, max_iter=1)
            eu += o.F.l2_error_sq(o.u, o.vdofs, tg.components)
        errs.append(np.sqrt(dt * eu))
    assert errs[1] < errs[0] / 4


@pytest.mark.parametrize("deg", [1, 2])
def test_assembly_with_bcs_strategies_agree(deg):
    """demo/assembly_bcs.py:132-234: the Oasis strategy (convection assembled, scaled and combined with M and K; RHS by
    mat-vec; matrix re-scaled `A <- -A + 2M/dt`; Dirichlet rows -> identity; `set_bc` on the RHS) and the direct one
    (LHS combined with the opposite signs, RHS integrated from the action) give the same vector and the same matrix --
    the reference raises RuntimeError otherwise (:224-234).  Here with the oracle's operators: it is the identity the
    product's `assemble_first` (GPU test `test_assemble_first_and_tentative_rhs`) is built on."""
    from oracle.ipcs_oracle import simplex_quadrature, tabulate, zero_rows

    from oasisx_b200 import fem, mesh as bmesh
    from problems import boundary_facets

    dt, nu = 0.5, 0.3
    msh = bmesh.create_unit_cube(None, 4, 3, 3)
    o = make_oracle(msh, deg, TaylorGreen(nu, 3), dt)
    V = fem.functionspace(msh, ("Lagrange", deg))
    bdofs = fem.locate_dofs_topological(V, 2, boundary_facets(msh))
    g = 2 * np.sin(o.xV[:, 0]) + 3 + 2 * o.xV[:, 1]                       # :50-51
    u1 = np.sin(o.xV[:, 0]) * np.cos(o.xV[:, 1])                          # :68
    uab = [o.xV[:, 0].copy() for _ in range(3)]                           # :76-77
    C = o.F.convection(uab)
    # Oasis approach (:132-167)
    A = -0.5 * C + o.M / dt - 0.5 * nu * o.K
    b = A @ u1
    b[bdofs] = g[bdofs]
    A = zero_rows((-A + (2.0 / dt) * o.M).tocsr(), bdofs, 1.0)
    # direct approach (:176-203)
    Ax = zero_rows((0.5 * C + o.M / dt + 0.5 * nu * o.K).tocsr(), bdofs, 1.0)
    F = o.F
    pts, w = simplex_quadrature(3, 2 * deg + 2)
    phi, dphi = tabulate(3, deg, pts)
    gp = np.einsum("cdk,qjd->cqjk", F.g.Kinv, dphi)
    uq = np.einsum("qa,ca->cq", phi, u1[o.vdofs])
    gu = np.einsum("cqjk,cj->cqk", gp, u1[o.vdofs])
    aq = np.stack([np.einsum("qa,ca->cq", phi, a[o.vdofs]) for a in uab], axis=2)
    be = np.einsum("c,q,cq,qi->ci", F.g.detJ, w, uq / dt - 0.5 * np.einsum("cqk,cqk->cq", aq, gu), phi) - 0.5 * nu * np.einsum(
        "c,q,cqk,cqik->ci", F.g.detJ, w, gu, gp)
    bx = np.zeros(o.nV)
    np.add.at(bx, o.vdofs.ravel(), be.ravel())
    bx[bdofs] = g[bdofs]
    assert np.allclose(bx, b)                                             # :224
    np.testing.assert_allclose(bx, b, rtol=1e-11, atol=1e-12)
    D = (A - Ax).tocsr()
    assert np.allclose(D.data, 0)                                         # :231
    assert abs(D).max() <= 1e-12 * abs(Ax).max()
    # Dirichlet rows are identity rows in both
    rows = Ax[bdofs].tocoo()  # (zeroRowsLocal keeps the pattern: stored zeros)
    live = rows.data != 0
    assert np.count_nonzero(live) == len(bdofs) and np.all(rows.data[live] == 1.0) and np.array_equal(rows.col[live], bdofs[rows.row[live]])
