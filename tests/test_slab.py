"""Slab-local box provider (oasisx_b200/slab.py): every rank builds only its own slab of the mesh; the result must be
the LocalProblem that cutting a replicated global mesh gives (oasisx_b200/partition.py) -- array for array -- and the
entity / dof searches the boundary conditions use must find the same dofs.  CPU only."""
import numpy as np
import pytest

from oasisx_b200 import fem, mesh as bmesh, multigrid, partition as part, slab


class FakeComm:
    def __init__(self, rank, size):
        self.rank, self.size = rank, size


def _meshes(gdim, shape, nranks, order="class"):
    pts = [[0.0, -1.0, 0.5][:gdim], [1.0, 2.0, 3.0][:gdim]]
    make = bmesh.create_box if gdim == 3 else bmesh.create_rectangle
    g = make(None, pts, list(shape))
    g._dof_order = order
    locs = []
    for r in range(nranks):
        m = make(FakeComm(r, nranks), pts, list(shape))
        assert isinstance(m, slab.SlabMesh)
        m._dof_order = order
        locs.append(m)
    return g, locs


CASES = [(3, (3, 2, 4), 2, 2, "class"), (3, (3, 2, 5), 3, 2, "class"), (3, (2, 3, 8), 8, 2, "class"),
         (3, (4, 3, 7), 4, 2, "generic"), (3, (3, 3, 6), 3, 1, "class"), (2, (5, 6), 2, 2, "class"),
         (2, (4, 9), 4, 2, "class"), (2, (4, 7), 3, 1, "class")]


@pytest.mark.parametrize("gdim,shape,nranks,deg_u,order", CASES)
def test_slab_local_problem_equals_partition_of_the_global_mesh(gdim, shape, nranks, deg_u, order):
    g, locs = _meshes(gdim, shape, nranks, order)
    gV, gQ = fem.functionspace(g, ("Lagrange", deg_u)), fem.functionspace(g, ("Lagrange", 1))
    total_owned = 0
    for r, m in enumerate(locs):
        ref = part.partition(g, gV, gQ, nranks, r)
        lp, V, Q = slab.local_problem(m, deg_u, 1)
        assert m.num_cells_global == g.num_cells
        assert np.array_equal(lp.cells, ref.cells) and lp.n_cells_owned == ref.n_cells_owned
        assert np.array_equal(m.geometry.x[lp.cell_nodes], g.geometry.x[ref.cell_nodes])  # bitwise the same coordinates
        for a, b in ((lp.V, ref.V), (lp.Q, ref.Q)):
            assert (a.n_owned, a.n_ghost, a.n_global) == (b.n_owned, b.n_ghost, b.n_global)
            for f in ("l2g", "cell_dofs", "x"):
                assert np.array_equal(getattr(a, f), getattr(b, f)), f
            for f in ("neighbors", "send_off", "send_idx", "recv_off"):
                assert np.array_equal(getattr(a.halo, f), getattr(b.halo, f)), f
            probe = np.concatenate([b.l2g[::3], [0, b.n_global - 1]])
            assert np.array_equal(a.g2l[probe], b.g2l[probe])
        total_owned += lp.V.n_owned
        # the spaces behave like fem.LocalFunctionSpace
        assert V.num_dofs == ref.V.n_local and V.dofmap.index_map.size_local == ref.V.n_owned
        assert np.array_equal(V.dofmap.index_map.ghosts, ref.V.l2g[ref.V.n_owned:])
    assert total_owned == gV.num_dofs


@pytest.mark.parametrize("gdim,shape,nranks", [(3, (3, 4, 6), 3), (2, (6, 8), 4)])
def test_boundary_searches_on_a_slab_find_the_global_boundary(gdim, shape, nranks):
    g, locs = _meshes(gdim, shape, nranks)
    gV = fem.functionspace(g, ("Lagrange", 2))
    gQ = fem.functionspace(g, ("Lagrange", 1))
    fdim = gdim - 1
    gfac = bmesh.exterior_facet_indices(g.topology)
    gdofs = fem.locate_dofs_topological(gV, fdim, gfac)
    top = lambda x: np.isclose(x[gdim - 1], [2.0, 3.0][gdim - 2])
    g_top = fem.locate_dofs_topological(gQ, fdim, bmesh.locate_entities_boundary(g, fdim, top))
    for r, m in enumerate(locs):
        lp, V, Q = slab.local_problem(m, 2, 1)
        fac = bmesh.exterior_facet_indices(m.topology)
        # no facet of a cut plane, every facet on the global boundary
        xf = m.geometry.x[m.topology.entities(fdim)[fac]]
        p0, p1 = m._box
        on = np.zeros(len(fac), dtype=bool)
        for a in range(gdim):
            on |= np.isclose(xf[:, :, a], p0[a]).all(axis=1) | np.isclose(xf[:, :, a], p1[a]).all(axis=1)
        assert on.all()
        dofs = fem.locate_dofs_topological(V, fdim, fac)
        want = lp.V.g2l[gdofs]
        assert np.array_equal(dofs, np.sort(want[want >= 0]))
        # a geometrically marked part of the boundary (only the last rank sees the top), pressure space
        ft = bmesh.locate_entities_boundary(m, fdim, top)
        assert (len(ft) > 0) == (r == nranks - 1)
        dq = fem.locate_dofs_topological(Q, fdim, ft)
        want = lp.Q.g2l[g_top]
        assert np.array_equal(dq, np.sort(want[want >= 0]))
        # geometrical search: local coordinates
        dg = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[0], 0.0))
        want = lp.V.g2l[fem.locate_dofs_geometrical(gV, lambda x: np.isclose(x[0], 0.0))]
        assert np.array_equal(dg, np.sort(want[want >= 0]))


def test_first_prolongation_restricted_to_owned_rows():
    fine, coarse = (4, 6, 8), (2, 3, 4)
    P = multigrid.lattice_prolongation(fine, coarse)
    rng = np.random.default_rng(0)
    idx = np.stack([rng.integers(0, s + 1, 40) for s in fine], axis=1)
    Pr = multigrid.lattice_prolongation(fine, coarse, idx)
    rows = multigrid._node_ids(idx.copy(), fine)
    assert abs(P[rows, :] - Pr).max() == 0


def test_global_mesh_switch(monkeypatch):
    monkeypatch.setenv("B200_GLOBAL_MESH", "1")
    m = bmesh.create_box(FakeComm(1, 2), [[0, 0, 0], [1, 1, 1]], [2, 2, 4])
    assert not isinstance(m, slab.SlabMesh) and m.num_cells == 6 * 16


def test_too_many_ranks_is_an_error():
    with pytest.raises(ValueError):
        bmesh.create_box(FakeComm(3, 4), [[0, 0, 0], [1, 1, 1]], [2, 2, 2])
