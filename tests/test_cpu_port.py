"""The C++/OpenMP CPU restatement (oracle/ipcs_cpu.cpp, used as bench.py's CPU baseline) against the
numpy/SuperLU oracle: patterns bit-exact, matrices <= 1e-12, fields after several steps <= 1e-8."""
import numpy as np
import pytest
import scipy.sparse as sp

from oasisx_b200 import fem
from oracle import ipcs_cpu as cpu
from problems import TaylorGreen, boundary_facets, make_mesh, make_oracle, relerr, vscale


def make_cpu(msh, deg_u, tg, dt, **kw):
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


@pytest.mark.parametrize("gdim,N,deg", [(2, 8, 2), (3, 4, 2), (2, 8, 1)])
def test_cpu_port_matches_numpy_oracle(gdim, N, deg):
    dt, nu = 0.005, 0.01
    msh = make_mesh(gdim, N)
    tg, tg2 = TaylorGreen(nu, gdim), TaylorGreen(nu, gdim)
    c = make_cpu(msh, deg, tg, dt, rtol=1e-13)
    o = make_oracle(msh, deg, tg2, dt)
    V, Q = fem.functionspace(msh, ("Lagrange", deg)), fem.functionspace(msh, ("Lagrange", 1))
    for which, (R, Cc) in enumerate([(V, V), (V, Q), (Q, V), (Q, Q)]):
        ip, ix = fem.build_csr_pattern(R.dofmap.list, Cc.dofmap.list, R.num_dofs, Cc.num_dofs)
        cip, cix = c.pattern(which, R.num_dofs)
        np.testing.assert_array_equal(cip, ip)
        np.testing.assert_array_equal(cix, ix)
    ip, ix = c.pattern(0, V.num_dofs)
    for which, ref in ((0, o.M), (1, o.K)):
        A = sp.csr_matrix((c.matrix(which, 0, len(ix)), ix, ip), shape=ref.shape)
        assert abs(A - ref).max() <= 1e-12 * abs(ref).max()
    ipq, ixq = c.pattern(1, V.num_dofs)
    for k in range(gdim):
        A = sp.csr_matrix((c.matrix(4, k, len(ixq)), ixq, ipq), shape=o.P[k].shape)
        assert abs(A - o.P[k]).max() <= 1e-12 * abs(o.P[k]).max()
    for t in (tg, tg2):
        t.t_u, t.t_p = 0.0, -dt / 2
    for n in range(3):
        for t in (tg, tg2):
            t.t_u += dt
            t.t_p += dt
        c.solve(dt, nu)
        o.solve(dt, nu, max_iter=1)
        for i in range(gdim):
            assert relerr(c.get(cpu.U, i), o.u[i], vscale(o.u)) <= 1e-8
        assert relerr(c.get(cpu.P), o.p) <= 1e-8
    assert (c.its > 0).all()


@pytest.mark.parametrize("order", [1, 2])
def test_cpu_port_bench_options_do_not_change_the_solution(order):
    """block-relative tolerance + extrapolated guesses (the bench settings; order 2 = quadratic history
    extrapolation, b200_guess=extrapolate2 on the GPU arm): same fields as the oracle."""
    dt, nu = 0.005, 0.01
    msh = make_mesh(3, 4)
    tg, tg2 = TaylorGreen(nu, 3), TaylorGreen(nu, 3)
    c = make_cpu(msh, 2, tg, dt, rtol=1e-12, nonzero_guess=True, block_rtol=True, extrapolate=order)
    o = make_oracle(msh, 2, tg2, dt)
    for t in (tg, tg2):
        t.t_u, t.t_p = 0.0, -dt / 2
    for n in range(5):
        for t in (tg, tg2):
            t.t_u += dt
            t.t_p += dt
        c.solve(dt, nu)
        o.solve(dt, nu, max_iter=1)
        for i in range(3):
            assert relerr(c.get(cpu.U, i), o.u[i], vscale(o.u)) <= 1e-8
        assert relerr(c.get(cpu.P), o.p) <= 1e-8


def test_cpu_port_lid_driven_cavity_matches_oracle():
    """bench.py's cavity CPU sample (merged wall + lid dofs, BiCGStab restarting on the start-up breakdown) against
    the LU oracle with the two constant DirichletBCs of the GPU test."""
    import bench
    from oasisx_b200 import mesh as bmesh
    from oracle.ipcs_oracle import OracleIPCS

    N, dt, nu = 4, 0.01, 0.01
    msh = bmesh.create_unit_cube(None, N, N, N)
    V, Q = fem.functionspace(msh, ("Lagrange", 2)), fem.functionspace(msh, ("Lagrange", 1))
    lid, walls = bench._cavity_markers()
    dw, dl = fem.locate_dofs_geometrical(V, walls), fem.locate_dofs_geometrical(V, lid)
    assert len(np.intersect1d(dw, dl)) == 0
    bd = fem.locate_dofs_geometrical(V, lambda x: lid(x) | walls(x))
    vals = [lambda x: np.where(lid(x), 1.0, 0.0), lambda x: np.zeros_like(x[0]), lambda x: np.zeros_like(x[0])]
    c = cpu.CpuIPCS(msh.geometry.x, msh.geometry.dofmap, 3, V.dofmap.list, Q.dofmap.list, V.tabulate_dof_coordinates(),
                    Q.tabulate_dof_coordinates(), 2, bcs_u=[[(bd, f)] for f in vals], rtol=1e-12, nonzero_guess=True,
                    block_rtol=True, extrapolate=2)
    o = OracleIPCS(msh.geometry.x, msh.geometry.dofmap, 3, V.dofmap.list, Q.dofmap.list, V.tabulate_dof_coordinates(),
                   Q.tabulate_dof_coordinates(), 2,
                   bcs_u=[[(dw, 0.0), (dl, 1.0)], [(dw, 0.0), (dl, 0.0)], [(dw, 0.0), (dl, 0.0)]])
    for n in range(4):
        c.solve(dt, nu)
        o.solve(dt, nu, max_iter=1)
        for i in range(3):
            assert relerr(c.get(cpu.U, i), o.u[i], vscale(o.u)) <= 1e-8, (n, i)
        assert relerr(c.get(cpu.P), o.p) <= 1e-8, n


@pytest.mark.parametrize("gdim,N", [(3, 8), (2, 16)])
def test_cpu_port_pressure_multigrid_matches_oracle(gdim, N):
    """The CPU arm's pc_type=mg (V(1,1), exact dense coarse solve: the GPU arm's algorithm) reproduces the oracle's
    direct pressure solve and needs far fewer iterations than Jacobi-PCG."""
    dt, nu = 0.005, 0.01
    msh = make_mesh(gdim, N)
    tg, tg2, tg3 = TaylorGreen(nu, gdim), TaylorGreen(nu, gdim), TaylorGreen(nu, gdim)
    c = make_cpu(msh, 2, tg, dt, rtol=1e-12, nonzero_guess=True, block_rtol=True, extrapolate=2)
    assert c.attach_pressure_multigrid(msh) >= 1
    cj = make_cpu(msh, 2, tg3, dt, rtol=1e-12, nonzero_guess=True, block_rtol=True, extrapolate=2)
    o = make_oracle(msh, 2, tg2, dt)
    for t in (tg, tg2, tg3):
        t.t_u, t.t_p = 0.0, -dt / 2
    for n in range(4):
        for t in (tg, tg2, tg3):
            t.t_u += dt
            t.t_p += dt
        c.solve(dt, nu)
        cj.solve(dt, nu)
        o.solve(dt, nu, max_iter=1)
        for i in range(gdim):
            assert relerr(c.get(cpu.U, i), o.u[i], vscale(o.u)) <= 1e-8
        assert relerr(c.get(cpu.P), o.p) <= 1e-8
        assert 0 < c.its[1] <= 25 and c.its[1] < cj.its[1], (c.its, cj.its)
