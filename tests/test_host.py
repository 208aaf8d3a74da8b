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
    assert ctypes.sizeof(_lib.Stats) == 8 * 4 + 8 + 5 * 8 + 16 + 16 + 24


def test_no_gpu_fails_loudly(lib):
    if lib.b2_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.B200Error):
        _lib.Context()
