"""Host provider: closed-form sizes, pattern properties, BC wrappers
(mirrors /root/reference/test/test_bcs.py with the DOLFINx comparison replaced by direct
evaluation), and the C-ABI export list."""
import ctypes
import os
import re

import numpy as np
import pytest

from oasisx_b200 import DirichletBC, LocatorMethod, fem, mesh as bmesh
from oasisx_b200 import _lib


@pytest.mark.parametrize("N", [3, 4, 5])
def test_box_counts_closed_form(N):
    msh = bmesh.create_box(None, [[-1, -1, -1], [1, 1, 1]], [N, N, N])
    V = fem.functionspace(msh, ("Lagrange", 2))
    Q = fem.functionspace(msh, ("Lagrange", 1))
    assert msh.num_cells == 6 * N**3
    assert msh.topology.num_entities(1) == 7 * N**3 + 9 * N**2 + 3 * N
    assert V.num_dofs == (2 * N + 1) ** 3 and Q.num_dofs == (N + 1) ** 3
    ip, ix = fem.build_csr_pattern(V.dofmap.list, V.dofmap.list, V.num_dofs, V.num_dofs)
    assert len(ix) == 230 * N**3 + 138 * N**2 + 24 * N + 1
    ip, ix = fem.build_csr_pattern(V.dofmap.list, Q.dofmap.list, V.num_dofs, Q.num_dofs)
    assert len(ix) == 65 * N**3 + 57 * N**2 + 15 * N + 1
    ip, ix = fem.build_csr_pattern(Q.dofmap.list, Q.dofmap.list, Q.num_dofs, Q.num_dofs)
    assert len(ix) == 15 * N**3 + 21 * N**2 + 9 * N + 1
    # all cells positively sized, total volume 8
    from oracle.ipcs_oracle import Geometry
    g = Geometry(msh.geometry.x, msh.geometry.dofmap, 3)
    np.testing.assert_allclose(g.detJ.sum() / 6, 8.0, rtol=1e-13)


def test_rectangle_counts():
    msh = bmesh.create_rectangle(None, [[-1, -1], [1, 1]], [64, 64])
    assert msh.num_cells == 8192 and msh.geometry.x.shape[0] == 4225
    assert msh.topology.num_entities(1) == 12416
    assert fem.functionspace(msh, ("Lagrange", 2)).num_dofs == 16641
    ext = bmesh.exterior_facet_indices(msh.topology)
    assert len(ext) == 4 * 64


def test_pattern_symmetric_and_sorted():
    msh = bmesh.create_unit_cube(None, 3, 2, 2)
    V = fem.functionspace(msh, ("Lagrange", 2))
    ip, ix = fem.build_csr_pattern(V.dofmap.list, V.dofmap.list, V.num_dofs, V.num_dofs)
    import scipy.sparse as sp
    A = sp.csr_matrix((np.ones(len(ix)), ix, ip), shape=(V.num_dofs,) * 2)
    assert (A != A.T).nnz == 0
    for r in range(V.num_dofs):
        row = ix[ip[r]:ip[r + 1]]
        assert np.all(np.diff(row) > 0) and r in row
    for cd in V.dofmap.list:  # every cell's dofs appear in each other's rows
        for r in cd:
            assert np.isin(cd, ix[ip[r]:ip[r + 1]]).all()


@pytest.mark.parametrize("P", [1, 2])
@pytest.mark.parametrize("dim", [0, 1])
def test_dirichlet_topological_matches_geometrical(P, dim):
    """test/test_bcs.py:58-97 in spirit: a time-dependent callable applied through the wrapper equals
    direct evaluation on the located dofs, for every update."""
    msh = bmesh.create_unit_square(None, 10, 10)
    locator = lambda x: np.isclose(x[0], 1)

    class TimeDependentBC:
        def __init__(self, t):
            self.t = t

        def eval(self, x):
            return np.sin(x[0]) + x[1] * self.t

    cond = TimeDependentBC(0.1)
    entities = bmesh.locate_entities(msh, dim, locator)
    value = np.int32(3)
    et = bmesh.meshtags(msh, dim, entities, np.full(len(entities), value, dtype=np.int32))
    bc = DirichletBC(cond.eval, LocatorMethod.TOPOLOGICAL, (et, value))
    V = fem.functionspace(msh, ("Lagrange", P))
    bc.create_bc(V)
    geo = fem.locate_dofs_geometrical(V, locator)
    if dim == 1:
        np.testing.assert_array_equal(np.sort(bc._dofs), geo)
    else:  # vertices only: the P2 edge midpoints are not in the closure of vertices
        assert np.isin(bc._dofs, geo).all()
    x = V.tabulate_dof_coordinates().T
    for t in [0.1, 0.2, 0.3]:
        cond.t = t
        bc.update_bc()
        u = fem.Function(V)
        bc.apply(u.x)
        expect = np.zeros(V.num_dofs)
        expect[bc._dofs] = (np.sin(x[0]) + x[1] * t)[bc._dofs]
        np.testing.assert_allclose(u.x.array, expect)


def test_constant_bc_reads_live_value():
    """test/test_bcs.py:100-134 in spirit."""
    msh = bmesh.create_unit_square(None, 10, 10)
    time = fem.Constant(msh, 1.0)
    bc = DirichletBC(time, LocatorMethod.GEOMETRICAL, lambda x: np.isclose(x[0], 1))
    V = fem.functionspace(msh, ("Lagrange", 2))
    bc.create_bc(V)
    for t in [0.1, 0.2, 0.3]:
        time.value = time.value + t
        u = fem.Function(V)
        bc.apply(u.x)
        assert np.allclose(u.x.array[bc._dofs], float(time.value)) and np.count_nonzero(u.x.array) == len(bc._dofs)


def test_abi_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "b200ipcs.h")).read()
    declared = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"b2_ctx", "b2_stats"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b2_abi_version() == 1
    assert ctypes.sizeof(_lib.Stats) == 8 * 4 + 8 + 5 * 8 + 16 + 16 + 24 + 8


def test_no_gpu_fails_loudly(lib):
    if lib.b2_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.B200Error):
        _lib.Context()


def test_taylor_green_spatial_cache_is_bitwise_and_keyed_by_array():
    """tests/problems.py caches the spatial factor of the separable exact solution per coordinate ARRAY (the boundary
    conditions pass the same array every step): same bits as the direct formula, and a different array is recomputed."""
    from problems import TaylorGreen

    nu = 0.01
    tg = TaylorGreen(nu, 3)
    rng = np.random.default_rng(3)
    x, y = rng.uniform(-1, 1, (3, 257)), rng.uniform(-1, 1, (3, 257))
    for t in (0.0, 0.3, 0.7):
        tg.t_u, tg.t_p = t, t - 0.0025
        for arr in (x, y, x):
            ex = -np.cos(np.pi * arr[0]) * np.sin(np.pi * arr[1]) * np.exp(-2.0 * nu * np.pi**2 * tg.t_u)
            ey = np.cos(np.pi * arr[1]) * np.sin(np.pi * arr[0]) * np.exp(-2.0 * nu * np.pi**2 * tg.t_u)
            ep = -0.25 * (np.cos(2 * np.pi * arr[0]) + np.cos(2 * np.pi * arr[1])) * np.exp(-4 * nu * np.pi**2 * tg.t_p)
            assert (tg.eval_x(arr) == ex).all() and (tg.eval_y(arr) == ey).all() and (tg.eval_p(arr) == ep).all()
            assert (tg.eval_z(arr) == 0).all()


def test_step_byte_accounting_is_consistent():
    """bench.step_algorithmic_bytes: stage sums add up and every extra Krylov iteration costs what SURVEY.md 8(d) says."""
    import bench

    n2, n1, nnz22, nnz21, nnz11 = 7189057, 912673, 204763393, 58034593, 13465441
    spmm = 12.0 * nnz22 + 4.0 * (n2 + 1) + 8.0 * 3 * 2 * n2
    spmv_q = 12.0 * nnz11 + 4.0 * (n1 + 1) + 16.0 * n1
    a = bench.step_algorithmic_bytes(n2, n1, nnz22, nnz21, nnz11, 3, (3, 9, 4), 13.5e9, spmm, spmv_q)
    b = bench.step_algorithmic_bytes(n2, n1, nnz22, nnz21, nnz11, 3, (4, 9, 4), 13.5e9, spmm, spmv_q)
    c = bench.step_algorithmic_bytes(n2, n1, nnz22, nnz21, nnz11, 3, (3, 9, 5), 13.5e9, spmm, spmv_q)
    assert abs(a["total"] - (a["assemble_first"] + a["tentative"] + a["pressure"] + a["update"]) - (2 * 24.0 * n2 + 16.0 * n1)) < 1.0
    assert abs((b["tentative"] - a["tentative"]) - (2 * spmm + 14 * 24.0 * n2)) < 1.0  # one BiCGStab iteration: 2 SpMM + 14 vector passes
    assert abs((c["update"] - a["update"]) - (spmm + 10 * 24.0 * n2)) < 1.0            # one CG iteration: 1 SpMM + 10 vector passes
    assert 70e9 < a["total"] < 100e9


def test_write_vtu_quadratic_cells(tmp_path):
    """State export (VTXWriter stand-in, demo/taylor_green.py:183-184): quadratic simplices in VTK node order."""
    from oasisx_b200 import mesh as bmesh
    from oasisx_b200.io import write_vtu

    for msh, ctype, npc in ((bmesh.create_unit_square(None, 3, 2), 22, 6), (bmesh.create_unit_cube(None, 2, 2, 2), 24, 10)):
        V = fem.functionspace(msh, ("Lagrange", 2))
        x = V.tabulate_dof_coordinates()
        path = tmp_path / f"m{ctype}.vtu"
        write_vtu(str(path), V, {"f": x[:, 0] + 2 * x[:, 1]})
        txt = path.read_text()
        assert f'NumberOfCells="{msh.num_cells}"' in txt and f'NumberOfPoints="{V.num_dofs}"' in txt
        conn = txt.split('Name="connectivity" format="ascii">\n')[1].split("</DataArray>")[0].split()
        assert len(conn) == msh.num_cells * npc
        # VTK quadratic simplex: node 3+k of a triangle / 4+k of a tetrahedron is the midpoint of its edge
        c = np.array(conn[:npc], dtype=int)
        nv = 3 if npc == 6 else 4
        edges = [(0, 1), (1, 2), (2, 0)] if npc == 6 else [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]
        for k, (a, b) in enumerate(edges):
            assert np.allclose(x[c[nv + k]], 0.5 * (x[c[a]] + x[c[b]]))


def test_host_comm_three_ranks_no_pickle(tmp_path):
    """The rendezvous channel carries JSON behind an HMAC handshake (advisor finding: it used pickle)."""
    import subprocess
    import sys

    import oasisx_b200.comm as comm_mod

    assert "pickle" not in open(comm_mod.__file__).read().replace("no pickle", "")
    script = tmp_path / "c.py"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script.write_text(
        "import sys\n"
        f"sys.path.insert(0, {root!r})\n"
        "from oasisx_b200.comm import HostComm\n"
        "c = HostComm.from_env()\n"
        "assert c.bcast(bytes(range(128)) if c.rank == 0 else None) == bytes(range(128))\n"
        "assert c.allreduce(c.rank + 0.5) == sum(r + 0.5 for r in range(c.size))\n"
        "assert c.allreduce(c.rank, 'max') == c.size - 1\n"
        "assert c.allgather({'n': [1, 2], 'r': (c.rank, 2.5)})[c.rank]['r'] == (c.rank, 2.5)\n"
        "try:\n    c.allreduce(1, 'prod')\n    raise SystemExit('no raise')\nexcept ValueError:\n    pass\n"
        "c.Barrier()\nprint('COMM_OK')\n")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(os.environ, RANK=str(r), WORLD_SIZE="3", MASTER_PORT="29877",
                                                                       B2_COMM_SECRET="s3cret"), stdout=subprocess.PIPE, text=True)
             for r in range(3)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs) and all("COMM_OK" in o for o in outs)


@pytest.mark.parametrize("kind,shape", [("box", (5, 4, 3)), ("rect", (7, 5)), ("box", (2, 3, 9))])
def test_closed_form_p2_space_is_the_general_route_bit_for_bit(kind, shape):
    """Box/rectangle meshes number their P2 dofs from the lattice in closed form (fem._lattice_p2: no edge table, no
    sort); the general route (edge table + class-order sort) must give the same cell dofs, bitwise the same dof
    coordinates and the same closure dofs of facets, edges and vertices."""
    make = lambda: (bmesh.create_box(None, [[0, -1, 0.5], [1, 2, 3]], list(shape)) if kind == "box"
                    else bmesh.create_rectangle(None, [[0, -1], [1, 2.5]], list(shape)))
    fast_mesh, slow_mesh = make(), make()
    slow_mesh._canonical = False
    V, G = fem.functionspace(fast_mesh, ("Lagrange", 2)), fem.functionspace(slow_mesh, ("Lagrange", 2))
    assert getattr(V, "_lattice_ids", False) and not getattr(G, "_lattice_ids", False)
    assert np.array_equal(V.dofmap.list, G.dofmap.list)
    assert np.array_equal(V.tabulate_dof_coordinates(), G.tabulate_dof_coordinates())
    d = fast_mesh.topology.dim
    fac = bmesh.exterior_facet_indices(fast_mesh.topology)
    assert np.array_equal(fac, bmesh.exterior_facet_indices(slow_mesh.topology))
    for edim, ents in ((d - 1, fac), (1, np.arange(0, fast_mesh.topology.num_entities(1), 3)), (0, np.arange(0, len(fast_mesh.geometry.x), 2))):
        assert np.array_equal(V.entity_closure_dofs(edim, ents), G.entity_closure_dofs(edim, ents))
