"""A duck-typed stand-in for the few DOLFINx objects ``oasisx_b200.adapter`` reads (SURVEY.md Appendix D list), built
from the built-in provider + partitioner -- DOLFINx itself cannot be installed here.  It reproduces the PROPERTIES the
adapter must cope with: owned dofs of a rank form a contiguous global range, local numbering is owned-first, and the
ghost block is in arbitrary (shuffled) order, not grouped by owner."""
from __future__ import annotations

import types

import numpy as np

from oasisx_b200 import fem, partition as part


class _IndexMap:
    def __init__(self, size_local, ghosts, owners, size_global, lo):
        self.size_local, self.num_ghosts = int(size_local), len(ghosts)
        self.ghosts, self.owners = np.asarray(ghosts, np.int64), np.asarray(owners, np.int32)
        self.size_global, self.local_range = int(size_global), (int(lo), int(lo) + int(size_local))


class _DofMap:
    def __init__(self, cell_dofs, index_map):
        self.list, self.index_map, self.index_map_bs = np.ascontiguousarray(cell_dofs, np.int32), index_map, 1

    def cell_dofs(self, c):
        return self.list[c]


class FakeSpace:
    def __init__(self, mesh, degree, lsp, owner_of_global, rng):
        self.mesh = mesh
        self.element = types.SimpleNamespace(degree=degree)
        n_owned, n_ghost = lsp.n_owned, lsp.n_ghost
        # DOLFINx-style global numbering: rank r owns [off[r], off[r] + n_owned_r), in the provider's global order
        nranks = int(owner_of_global.max()) + 1
        counts = np.bincount(owner_of_global, minlength=nranks)
        off = np.concatenate([[0], np.cumsum(counts)])
        newg = np.empty(len(owner_of_global), dtype=np.int64)
        for r in range(nranks):
            mine = np.flatnonzero(owner_of_global == r)
            newg[mine] = off[r] + np.arange(len(mine))
        self._newg = newg
        shuffle = rng.permutation(n_ghost)                      # fake local ghost k = provider ghost shuffle[k]
        old_of_fake = np.concatenate([np.arange(n_owned), n_owned + shuffle])
        fake_of_old = np.empty_like(old_of_fake)
        fake_of_old[old_of_fake] = np.arange(len(old_of_fake))
        self._fake_of_old = fake_of_old
        gl = lsp.l2g[old_of_fake]
        self._x = lsp.x[old_of_fake]
        self.dofmap = _DofMap(fake_of_old[lsp.cell_dofs], _IndexMap(n_owned, newg[gl[n_owned:]], owner_of_global[gl[n_owned:]],
                                                                   len(owner_of_global), off[mesh.comm.rank]))

    def tabulate_dof_coordinates(self):
        return self._x


class FakeComm:
    """allgather over pre-collected per-rank values (the test drives all 'ranks' in one process, in two passes)."""

    def __init__(self, rank, size, board):
        self.rank, self.size, self._board = rank, size, board

    def allgather(self, value):
        if self.size == 1:
            return [value]
        slot = self._board.setdefault("calls", {}).setdefault(self._board["phase"], {})
        key = self._board.setdefault("n", {}).get(self.rank, 0)
        self._board["n"][self.rank] = key + 1
        slot.setdefault(key, {})[self.rank] = value
        known = self._board.get("known", {}).get(key)
        if known is None:
            raise _NeedOthers()
        return [known[r] for r in range(self.size)]


class _NeedOthers(Exception):
    pass


def make_fake(msh, deg_u, deg_p, nranks, rank, board=None, seed=0):
    """(fake dolfinx module, fake mesh, provider LocalProblem) for one rank of an `nranks`-way partition of `msh`."""
    V, Q = fem.functionspace(msh, ("Lagrange", deg_u)), fem.functionspace(msh, ("Lagrange", deg_p))
    lp = part.partition(msh, V, Q, nranks, rank) if nranks > 1 else None
    rng = np.random.default_rng(seed + rank)
    if lp is None:
        crank = np.zeros(msh.num_cells, np.int32)
        ownV, ownQ = np.zeros(V.num_dofs, np.int32), np.zeros(Q.num_dofs, np.int32)
        lp = part.partition(msh, V, Q, 1, 0)
    else:
        crank = part.cell_ranks(msh, nranks)
        ownV, ownQ = part._owners(V.dofmap.list, crank, V.num_dofs, nranks), part._owners(Q.dofmap.list, crank, Q.num_dofs, nranks)
    comm = FakeComm(rank, nranks, board if board is not None else {"phase": 0})
    tdim = msh.topology.dim
    def connectivity(d0, d1):
        # cell -> facet, local cells (owned then ghost) in reference-cell facet order; the facet ids are the provider's
        assert d0 == tdim and d1 == tdim - 1
        return types.SimpleNamespace(array=msh.topology.cell_entities(d1)[lp.cells].ravel())

    mesh = types.SimpleNamespace(
        comm=comm,
        geometry=types.SimpleNamespace(dim=msh.geometry.dim, x=msh.geometry.x, dofmap=lp.cell_nodes),
        topology=types.SimpleNamespace(dim=tdim, create_connectivity=lambda a, b: None, connectivity=connectivity,
                                       index_map=lambda d: types.SimpleNamespace(size_local=lp.n_cells_owned,
                                                                                 num_ghosts=len(lp.cells) - lp.n_cells_owned)))
    spaces = {deg_u: FakeSpace(mesh, deg_u, lp.V, ownV, rng), ("q", deg_p): FakeSpace(mesh, deg_p, lp.Q, ownQ, rng)}
    calls = {"n": 0}

    def functionspace(m, element):
        calls["n"] += 1
        return spaces[int(element[1])] if calls["n"] % 2 == 1 else spaces[("q", int(element[1]))]

    gV = V

    def locate_dofs_topological(Vf, edim, entities):
        # the provider locates GLOBAL dofs; hand back the fake LOCAL indices of those present on this rank
        g = (gV if Vf is spaces[deg_u] else Q).entity_closure_dofs(edim, np.asarray(entities))
        lsp = lp.V if Vf is spaces[deg_u] else lp.Q
        loc = lsp.g2l[g]
        return Vf._fake_of_old[loc[loc >= 0]].astype(np.int32)

    def fake_tags(tags):
        """MeshTags of the provider mesh re-attached to the fake mesh (same facet ids)."""
        return types.SimpleNamespace(topology=mesh.topology, dim=tags.dim, indices=tags.indices, values=tags.values, find=tags.find)

    mesh.fake_tags = fake_tags
    mod = types.SimpleNamespace(fem=types.SimpleNamespace(functionspace=functionspace, locate_dofs_topological=locate_dofs_topological))
    return mod, mesh, lp
