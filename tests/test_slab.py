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


def test_slab_halo_over_gloo_two_ranks(tmp_path):
    """world_size 2 on CPU (torch.distributed/gloo for the data, oasisx_b200.comm.HostComm as the mesh communicator):
    each rank builds ONLY its slab, the ranks cross-check their halo plans, exchange owner values with exactly the
    message pattern of the device halo (send_idx / recv_off) and all-reduce a dot product over the owned dofs -- the
    result is the global vector / the global dot product, which no rank ever assembled from a global mesh."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "slab2.py"
    script.write_text(
        "import os, sys\n"
        f"sys.path.insert(0, {root!r})\n"
        "import numpy as np, torch, torch.distributed as dist\n"
        "from oasisx_b200 import mesh as bmesh, partition as part, slab\n"
        "from oasisx_b200.comm import HostComm\n"
        "dist.init_process_group('gloo')\n"
        "r, n = dist.get_rank(), dist.get_world_size()\n"
        "os.environ['MASTER_PORT'] = str(int(os.environ['MASTER_PORT']) + 17)  # HostComm's own channel\n"
        "comm = HostComm.from_env()\n"
        "msh = bmesh.create_box(comm, [[-1, -1, -1], [1, 1, 1]], [5, 4, 6])\n"
        "assert isinstance(msh, slab.SlabMesh) and msh.num_cells < msh.num_cells_global\n"
        "lp, V, Q = slab.local_problem(msh, 2, 1)\n"
        "for name, s in (('V', lp.V), ('Q', lp.Q)):\n"
        "    part.check_halo_counts(comm, s.halo, name)\n"
        "    f = lambda x: np.sin(3 * x[:, 0]) + x[:, 1] * x[:, 2]  # a field known from the coordinates alone\n"
        "    v = torch.full((s.n_local,), float('nan'), dtype=torch.float64)\n"
        "    v[: s.n_owned] = torch.from_numpy(f(s.x[: s.n_owned]))\n"
        "    ops, bufs = [], []\n"
        "    for k, q in enumerate(s.halo.neighbors):\n"
        "        send = v[torch.from_numpy(s.halo.send_idx[s.halo.send_off[k]:s.halo.send_off[k+1]].astype(np.int64))].contiguous()\n"
        "        recv = torch.empty(int(s.halo.recv_off[k+1] - s.halo.recv_off[k]), dtype=torch.float64)\n"
        "        bufs.append((k, recv))\n"
        "        ops += [dist.P2POp(dist.isend, send, int(q)), dist.P2POp(dist.irecv, recv, int(q))]\n"
        "    for w in dist.batch_isend_irecv(ops): w.wait()\n"
        "    for k, recv in bufs:\n"
        "        v[s.n_owned + int(s.halo.recv_off[k]): s.n_owned + int(s.halo.recv_off[k+1])] = recv\n"
        "    assert np.array_equal(v.numpy(), f(s.x)), (r, name)\n"
        "    t = torch.tensor([float(s.n_owned), float(np.sum(s.l2g[: s.n_owned]))], dtype=torch.float64)\n"
        "    dist.all_reduce(t)\n"
        "    assert t[0].item() == s.n_global and t[1].item() == s.n_global * (s.n_global - 1) / 2, (r, name)  # every dof owned once\n"
        "comm.Barrier()\n"
        "dist.destroy_process_group()\n"
        "print('ok', r)\n"
    )
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29633", str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


@pytest.mark.parametrize("gdim,shape,nranks", [(2, (10, 10), 2), (3, (4, 4, 6), 3)])
def test_pressure_bc_on_a_slab_equals_the_partitioned_route(gdim, shape, nranks):
    """PressureBC.create_bcs on slab-local spaces: the same (local cell, local facet) pairs and the same pressure Dirichlet
    dofs as on the LocalFunctionSpaces of a partitioned replicated mesh (tests/test_partition.py covers that route)."""
    from oasisx_b200 import PressureBC

    g, locs = _meshes(gdim, shape, nranks)
    fdim = gdim - 1
    right = lambda x: np.isclose(x[0], 1.0)
    gV, gQ = fem.functionspace(g, ("Lagrange", 2)), fem.functionspace(g, ("Lagrange", 1))
    gf = bmesh.locate_entities_boundary(g, fdim, right)
    gtags = bmesh.meshtags(g, fdim, np.sort(gf), np.full(len(gf), 3, dtype=np.int32))
    for r, m in enumerate(locs):
        ref_lp = part.partition(g, gV, gQ, nranks, r)
        V, Q = fem.LocalFunctionSpace(gV, ref_lp.V, 1), fem.LocalFunctionSpace(gQ, ref_lp.Q, 1)
        V._local_cells = Q._local_cells = ref_lp.cells
        ref = PressureBC(lambda x: 2.0 + x[1], (gtags, 3))
        ref.create_bcs(V, Q)
        lp, Vs, Qs = slab.local_problem(m, 2, 1)
        lf = bmesh.locate_entities_boundary(m, fdim, right)
        ltags = bmesh.meshtags(m, fdim, np.sort(lf), np.full(len(lf), 3, dtype=np.int32))
        bc = PressureBC(lambda x: 2.0 + x[1], (ltags, 3))
        bc.create_bcs(Vs, Qs)
        assert sorted(zip(bc._facet_cells.tolist(), bc._facet_local.tolist())) == sorted(zip(ref._facet_cells.tolist(), ref._facet_local.tolist()))
        assert np.array_equal(np.sort(bc.bc.dofs), np.sort(ref.bc.dofs))
        assert np.array_equal(bc._h, ref._h)
