"""Brick SpMM (k_spmm_brick: x staged in shared memory, 16-bit positions) against the plain sliced-ELL kernel and the
oracle: same slots and the same order of FMAs per row, so single products are BITWISE equal; whole steps agree to
rounding (the fused dot products are summed in a different order)."""
import numpy as np
import pytest

from oasisx_b200 import _lib as L
from problems import TaylorGreen, TaylorGreenRot, make_mesh, make_oracle, make_solver, relerr, vscale

pytestmark = pytest.mark.gpu

KRYLOV = {
    "tentative": {"ksp_type": "bcgs", "pc_type": "jacobi", "ksp_rtol": 1e-12, "ksp_initial_guess_nonzero": True},
    "pressure": {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-12},
    "scalar": {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-12, "ksp_initial_guess_nonzero": True},
}


@pytest.mark.parametrize("gdim,N,order", [(3, 8, "class"), (3, 7, "generic"), (2, 40, "class"), (3, 33, "class")])
def test_brick_product_is_bitwise_the_plain_product(gdim, N, order):
    msh = make_mesh(gdim, N)
    msh._dof_order = order
    tg = TaylorGreen(0.01, gdim)
    s = make_solver(msh, 2, tg, 0.01, solver_options=KRYLOV, bricks=True)
    info = s._brick_info
    assert info and info["bricks"] >= 1 and info["max_gather"] <= 4352
    ctx = s._ctx
    rng = np.random.default_rng(5)
    tg.t_u, tg.t_p = 0.01, 0.005
    s.assemble_first(0.01, 0.01)  # A: non-symmetric values through the scatter table
    n = s._M.getSize()[1]
    for mat in (s._M, s._K, s._A):
        x = rng.standard_normal(n)
        yb, yb2, yp = (np.zeros(mat.getSize()[0]) for _ in range(3))
        ctx.set_tuning("spmm_brick", 1)
        mat.mult(x, yb)
        ctx.set_tuning("spmm_brick", 2)  # the pipelined kernel (TMA-fed rings)
        mat.mult(x, yb2)
        ctx.set_tuning("spmm_brick", 3)  # the same pipeline fed by 16-byte cp.async
        yb3 = np.zeros(mat.getSize()[0])
        mat.mult(x, yb3)
        ctx.set_tuning("spmm_brick", 0)
        mat.mult(x, yp)
        ctx.set_tuning("spmm_brick", 1)
        assert np.array_equal(yb, yp)
        assert np.array_equal(yb2, yp)
        assert np.array_equal(yb3, yp)
        ip, ix, v = mat.getValuesCSR()
        import scipy.sparse as sp
        ref = sp.csr_matrix((v, ix, ip), shape=mat.getSize()) @ x
        assert relerr(yb, ref) <= 1e-13


@pytest.mark.parametrize("gdim,N", [(3, 6), (2, 24)])
def test_steps_with_bricks_match_plain_kernel_and_oracle(gdim, N):
    dt, nu = 0.005, 0.01
    fields = []
    for brick in (2, 1, 0):
        msh = make_mesh(gdim, N)
        tg = TaylorGreenRot(nu) if gdim == 3 else TaylorGreen(nu, 2)
        s = make_solver(msh, 2, tg, dt, solver_options=KRYLOV, bricks=True)
        s._ctx.set_tuning("spmm_brick", brick)
        tg.t_u, tg.t_p = 0.0, -dt / 2
        its = []
        for _ in range(3):
            tg.t_u += dt
            tg.t_p += dt
            s.solve(dt, nu, max_iter=1)
            st = s._ctx.stats()
            its.append((tuple(st.its_tentative), st.its_pressure, tuple(st.its_update)))
        fields.append(([s._u[i].x.array_ro().copy() for i in range(gdim)], s._p.x.array_ro().copy(), its))
    (ub, pb, ib), (ub1, pb1, _), (up, pp, ip_) = fields
    sc = vscale(up)
    for i in range(gdim):
        assert relerr(ub[i], up[i], sc) <= 1e-11
        assert relerr(ub1[i], up[i], sc) <= 1e-11
    assert relerr(pb, pp) <= 1e-10 and relerr(pb1, pp) <= 1e-10
    # and against the LU oracle, at the tolerance of every other step test
    msh = make_mesh(gdim, N)
    tg = TaylorGreenRot(nu) if gdim == 3 else TaylorGreen(nu, 2)
    o = make_oracle(msh, 2, tg, dt)
    tg.t_u, tg.t_p = 0.0, -dt / 2
    for _ in range(3):
        tg.t_u += dt
        tg.t_p += dt
        o.solve(dt, nu, max_iter=1)
    for i in range(gdim):
        assert relerr(ub[i], o.u[i], vscale(o.u)) <= 1e-8
    assert relerr(pb, o.p) <= 1e-8
