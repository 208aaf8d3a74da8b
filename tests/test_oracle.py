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
