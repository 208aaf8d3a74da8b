"""Host-side mesh provider (numpy only).

Stands in for the parts of ``dolfinx.mesh`` the IPCS hot path consumes:
``create_rectangle`` / ``create_unit_square`` (triangles, "right" diagonal) and ``create_box`` /
``create_unit_cube`` (6 Kuhn tetrahedra per cube sharing the v0-v7 diagonal), exterior facets,
``locate_entities[_boundary]`` and ``meshtags`` as used by
``/root/reference/demo/taylor_green.py:126-140`` and
``/root/reference/test/test_tentative_velocity.py:90-128``.

Nothing downstream depends on the lattice structure: kernels only see ``geometry.x``,
``geometry.dofmap`` and the per-cell dof maps, exactly the arrays the DOLFINx adapter
(``oasisx_b200.adapter``) extracts from a real ``dolfinx.mesh.Mesh``.
"""
from __future__ import annotations

from enum import Enum

import numpy as np

__all__ = [
    "CellType",
    "Mesh",
    "MeshTags",
    "create_rectangle",
    "create_unit_square",
    "create_box",
    "create_unit_cube",
    "exterior_facet_indices",
    "locate_entities",
    "locate_entities_boundary",
    "meshtags",
]


class CellType(Enum):
    triangle = 2
    tetrahedron = 3


class _SerialComm:
    """Minimal stand-in for an MPI communicator (rank/size + allreduce on one rank)."""

    rank = 0
    size = 1

    def allreduce(self, value, op=None):
        return value

    def Barrier(self):
        pass

    def gather(self, value, root=0):
        return [value]


COMM_SELF = _SerialComm()

# local vertex pairs of the edges of a simplex, basix/UFC order
_EDGE_VERTS = {
    2: np.array([[1, 2], [0, 2], [0, 1]], dtype=np.int64),
    3: np.array([[2, 3], [1, 3], [1, 2], [0, 3], [0, 2], [0, 1]], dtype=np.int64),
}
# cells of one square / cube of the box providers as corner numbers (corner c has offsets (c & 1, c >> 1 & 1, c >> 2))
_BOX_CELLS = {
    2: [(0, 1, 3), (0, 2, 3)],
    3: [(0, 1, 3, 7), (0, 1, 7, 5), (0, 5, 7, 4), (0, 3, 2, 7), (0, 6, 4, 7), (0, 2, 6, 7)],
}
# local vertices of the facets (facet i is opposite vertex i)
_FACET_VERTS = {
    2: np.array([[1, 2], [0, 2], [0, 1]], dtype=np.int64),
    3: np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]], dtype=np.int64),
}


def _unique_rows(rows: np.ndarray, nv: int):
    """Unique rows of a (n, k<=3) array of sorted vertex ids < nv.  Returns (uniq, inverse)."""
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    k = rows.shape[1]
    if k == 1:
        u, inv = np.unique(rows[:, 0], return_inverse=True)
        return u[:, None], inv
    if k == 2 or nv**3 < 2**62:
        key = rows[:, 0]
        for c in range(1, k):
            key = key * nv + rows[:, c]
        u, inv = np.unique(key, return_inverse=True)
        out = np.empty((len(u), k), dtype=np.int64)
        for c in range(k - 1, -1, -1):
            out[:, c] = u % nv
            u = u // nv
        return out, inv
    # very large meshes: two-key lexsort
    k1 = rows[:, 0] * nv + rows[:, 1]
    order = np.lexsort((rows[:, 2], k1))
    s1, s2 = k1[order], rows[order, 2]
    new = np.ones(len(order), dtype=bool)
    new[1:] = (s1[1:] != s1[:-1]) | (s2[1:] != s2[:-1])
    ids = np.cumsum(new) - 1
    inv = np.empty(len(order), dtype=np.int64)
    inv[order] = ids
    return rows[order][new], inv


class Geometry:
    def __init__(self, x: np.ndarray, dofmap: np.ndarray, dim: int):
        self.x = x  # (n_nodes, 3) float64, padded with zeros like dolfinx
        self.dofmap = dofmap  # (n_cells, dim+1) int32
        self.dim = dim


class Topology:
    """Lazily computed entity numbering: vertices (0), edges (1), facets (dim-1), cells (dim)."""

    def __init__(self, mesh: "Mesh"):
        self._mesh = mesh
        self.dim = mesh.geometry.dim
        self._ent = {}  # entity dim -> (entities (n,k) vertex ids, cell_entities (n_cells, m))

    def _build(self, edim: int):
        if edim in self._ent:
            return self._ent[edim]
        cells = self._mesh.geometry.dofmap.astype(np.int64)
        nv = self._mesh.geometry.x.shape[0]
        d = self.dim
        if edim == 0:
            res = (np.arange(nv, dtype=np.int64)[:, None], cells)
        elif edim == d:
            res = (cells, np.arange(len(cells), dtype=np.int64)[:, None])
        else:
            loc = _EDGE_VERTS[d] if edim == 1 else _FACET_VERTS[d]
            ev = np.sort(cells[:, loc], axis=2).reshape(-1, loc.shape[1])
            uniq, inv = _unique_rows(ev, nv)
            res = (uniq, inv.reshape(len(cells), loc.shape[0]))
        self._ent[edim] = res
        return res

    def create_connectivity(self, d0: int, d1: int):
        self._build(d0)
        self._build(d1)

    def create_entities(self, edim: int):
        self._build(edim)

    def entities(self, edim: int) -> np.ndarray:
        """(n_entities, n_vertices_per_entity) vertex ids (sorted)."""
        return self._build(edim)[0]

    def cell_entities(self, edim: int) -> np.ndarray:
        """(n_cells, n_entities_per_cell) entity ids in basix/UFC local order."""
        return self._build(edim)[1]

    def num_entities(self, edim: int) -> int:
        return len(self._build(edim)[0])


class Mesh:
    def __init__(self, x: np.ndarray, cells: np.ndarray, gdim: int, comm=None):
        x = np.asarray(x, dtype=np.float64)
        if x.shape[1] < 3:
            x = np.hstack([x, np.zeros((x.shape[0], 3 - x.shape[1]))])
        self.geometry = Geometry(np.ascontiguousarray(x), np.ascontiguousarray(cells, dtype=np.int32), gdim)
        self.topology = Topology(self)
        self.comm = COMM_SELF if comm is None else comm

    @property
    def num_cells(self) -> int:
        return self.geometry.dofmap.shape[0]

    def cell_name(self) -> str:
        return "triangle" if self.geometry.dim == 2 else "tetrahedron"

    def h(self, dim: int, entities: np.ndarray) -> np.ndarray:
        """Cell diameter (longest edge), the quantity ``demo/taylor_green.py:217-224`` reduces."""
        assert dim == self.topology.dim
        c = self.geometry.dofmap[np.asarray(entities)]
        x = self.geometry.x[c]
        hmax = np.zeros(len(c))
        for a, b in _EDGE_VERTS[self.geometry.dim]:
            hmax = np.maximum(hmax, np.linalg.norm(x[:, a] - x[:, b], axis=1))
        return hmax


class MeshTags:
    """``dolfinx.mesh.MeshTags`` look-alike: sorted unique ``indices`` with ``values``."""

    def __init__(self, mesh: Mesh, dim: int, indices, values):
        self.mesh = mesh
        self.topology = mesh.topology
        self.dim = int(dim)
        self.indices = np.asarray(indices, dtype=np.int32)
        self.values = np.asarray(values)

    def find(self, value) -> np.ndarray:
        return self.indices[self.values == value]


def meshtags(mesh: Mesh, dim: int, entities, values) -> MeshTags:
    return MeshTags(mesh, dim, entities, values)


def _distributed(comm) -> bool:
    """Several ranks: every rank builds only its slab (oasisx_b200.slab), as DOLFINx's create_box(MPI.COMM_WORLD, ...)
    does [ext]; B200_GLOBAL_MESH=1 keeps the replicated mesh that oasisx_b200.partition cuts afterwards."""
    import os

    return comm is not None and int(getattr(comm, "size", 1)) > 1 and os.environ.get("B200_GLOBAL_MESH", "0") != "1"


def create_rectangle(comm, points, n, cell_type=CellType.triangle) -> Mesh:
    """Rectangle split into nx*ny squares, each into 2 triangles along the "right" diagonal
    (v0,v1,v3),(v0,v2,v3) -- the DOLFINx default [ext]."""
    if _distributed(comm):
        from .slab import create_slab_mesh

        return create_slab_mesh(comm, points, n, 2)
    (x0, y0), (x1, y1) = np.asarray(points, dtype=np.float64)[:, :2]
    nx, ny = int(n[0]), int(n[1])
    xs = np.linspace(x0, x1, nx + 1)
    ys = np.linspace(y0, y1, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")  # row = iy
    x = np.stack([X.ravel(), Y.ravel()], axis=1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (iy * (nx + 1) + ix).ravel()
    v1, v2, v3 = v0 + 1, v0 + nx + 1, v0 + nx + 2
    cells = np.empty((nx * ny, 2, 3), dtype=np.int64)
    cells[:, 0] = np.stack([v0, v1, v3], axis=1)
    cells[:, 1] = np.stack([v0, v2, v3], axis=1)
    msh = Mesh(x, cells.reshape(-1, 3), 2, comm)
    msh._lattice = (np.array([x0, y0, 0.0]), np.array([(x1 - x0) / nx, (y1 - y0) / ny, 1.0]))
    msh._shape = (nx, ny)
    msh._box = (np.array([x0, y0]), np.array([x1, y1]))
    msh._canonical = True  # nodes lexicographic, cells square by square in the order above (fem._lattice_p2 relies on it)
    return msh


def create_unit_square(comm, nx, ny, cell_type=CellType.triangle) -> Mesh:
    return create_rectangle(comm, [[0.0, 0.0], [1.0, 1.0]], [nx, ny], cell_type)


def create_box(comm, points, n, cell_type=CellType.tetrahedron) -> Mesh:
    """Box split into nx*ny*nz cubes, each into 6 tetrahedra around the v0-v7 diagonal."""
    if _distributed(comm):
        from .slab import create_slab_mesh

        return create_slab_mesh(comm, points, n, 3)
    p0, p1 = np.asarray(points, dtype=np.float64)
    nx, ny, nz = (int(v) for v in n)
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    zs = np.linspace(p0[2], p1[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    x = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    v0 = (iz * sz + iy * sy + ix).ravel()
    v = [v0, v0 + sx, v0 + sy, v0 + sx + sy, v0 + sz, v0 + sx + sz, v0 + sy + sz, v0 + sx + sy + sz]
    tets = _BOX_CELLS[3]
    cells = np.empty((len(v0), 6, 4), dtype=np.int64)
    for t, tet in enumerate(tets):
        cells[:, t] = np.stack([v[a] for a in tet], axis=1)
    msh = Mesh(x, cells.reshape(-1, 4), 3, comm)
    msh._lattice = (p0.copy(), (p1 - p0) / np.array([nx, ny, nz], dtype=np.float64))
    msh._shape = (nx, ny, nz)
    msh._box = (p0.copy(), p1.copy())
    msh._canonical = True  # nodes lexicographic, cells cube by cube in the order above (fem._lattice_p2 relies on it)
    return msh


def create_unit_cube(comm, nx, ny, nz, cell_type=CellType.tetrahedron) -> Mesh:
    return create_box(comm, [[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]], [nx, ny, nz], cell_type)


def exterior_facet_indices(topology: Topology) -> np.ndarray:
    """Facets attached to exactly one cell (sorted)."""
    fdim = topology.dim - 1
    cf = topology.cell_entities(fdim)
    counts = np.bincount(cf.ravel(), minlength=topology.num_entities(fdim))
    ext = np.flatnonzero(counts == 1)
    msh = topology._mesh
    if hasattr(msh, "is_global_boundary"):  # slab-local mesh: the cut planes between the ranks' slabs are not exterior
        ext = ext[msh.is_global_boundary(topology.entities(fdim)[ext])]
    return ext.astype(np.int32)


def _entity_marker(mesh: Mesh, edim: int, marker, candidates=None) -> np.ndarray:
    ents = mesh.topology.entities(edim)
    if candidates is not None:
        ents = ents[candidates]
    x = mesh.geometry.x
    ok = np.ones(len(ents), dtype=bool)
    marked_v = np.asarray(marker(x.T), dtype=bool)
    for c in range(ents.shape[1]):
        ok &= marked_v[ents[:, c]]
    idx = np.flatnonzero(ok)
    if candidates is not None:
        idx = np.asarray(candidates)[idx]
    return idx.astype(np.int32)


def locate_entities(mesh: Mesh, dim: int, marker) -> np.ndarray:
    """Entities all of whose vertices satisfy ``marker(x)`` (x is (3, n))."""
    return _entity_marker(mesh, dim, marker)


def locate_entities_boundary(mesh: Mesh, dim: int, marker) -> np.ndarray:
    """As :func:`locate_entities`, restricted to entities on the boundary."""
    top = mesh.topology
    fdim = top.dim - 1
    ext = exterior_facet_indices(top)
    if dim == fdim:
        return _entity_marker(mesh, dim, marker, ext)
    # vertices / edges of exterior facets
    bverts = np.unique(top.entities(fdim)[ext].ravel())
    on_b = np.zeros(mesh.geometry.x.shape[0], dtype=bool)
    on_b[bverts] = True
    ents = top.entities(dim)
    cand = np.flatnonzero(on_b[ents].all(axis=1))
    return _entity_marker(mesh, dim, marker, cand)
