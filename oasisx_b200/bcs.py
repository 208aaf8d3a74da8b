"""Boundary-condition wrappers with the surface of ``/root/reference/src/oasisx/bcs.py``.

``DirichletBC`` (``bcs.py:36-139``) keeps the constructor, ``set_dofs``, ``create_bc``,
``update_bc`` and ``apply``; values are held only on the BC dofs, because that is all
``set_bc`` ever reads (the reference re-interpolates every cell each step).
``PressureBC`` (``bcs.py:142-268``) keeps ``create_bcs``, ``update_bc``, ``bc`` and ``rhs``.
"""
from __future__ import annotations

from enum import Enum
from typing import Callable

import numpy as np

from . import fem as _fem
from .mesh import MeshTags

__all__ = ["DirichletBC", "PressureBC", "LocatorMethod"]


class LocatorMethod(Enum):
    """Search methods for Dirichlet BCs (``bcs.py:23-33``)."""

    GEOMETRICAL = 1
    TOPOLOGICAL = 2


class _DofBC:
    """What ``dolfinx.fem.dirichletbc`` hands to ``set_bc``: dofs and the values on them."""

    def __init__(self, dofs: np.ndarray, values: np.ndarray):
        self.dofs = dofs
        self.values = values

    def dof_indices(self):
        return self.dofs, len(self.dofs)


class DirichletBC:
    """Dirichlet condition located topologically or geometrically (``bcs.py:36-101``).

    Args:
        value: a float, a :class:`oasisx_b200.fem.Constant` or a callable ``f(x)``.
        method: :class:`LocatorMethod`.
        marker: ``(MeshTags, value)`` for TOPOLOGICAL, a callable ``x -> bool mask`` for GEOMETRICAL.
    """

    def __init__(self, value, method: LocatorMethod, marker):
        if method == LocatorMethod.GEOMETRICAL:
            self._method = method
            self._locator = marker
        elif method == LocatorMethod.TOPOLOGICAL:
            self._method = method
            self._entities = marker[0].find(marker[1])
            self._e_dim = marker[0].dim
        else:
            raise ValueError(method)
        self._value = value
        self._version = 0

    def set_dofs(self, dofs: np.ndarray):
        self._dofs = np.asarray(dofs, dtype=np.int32)

    def _locate_dofs(self, V: _fem.FunctionSpace):
        if self._method == LocatorMethod.GEOMETRICAL:
            self._dofs = _fem.locate_dofs_geometrical(V, self._locator)
        else:
            V.mesh.topology.create_connectivity(self._e_dim, V.mesh.topology.dim)
            self._dofs = _fem.locate_dofs_topological(V, self._e_dim, self._entities)

    def create_bc(self, V: _fem.FunctionSpace):
        """``bcs.py:116-126``."""
        if not hasattr(self, "_dofs"):
            self._locate_dofs(V)
        self._V = V
        self._xT = np.ascontiguousarray(V.tabulate_dof_coordinates()[self._dofs].T)
        self._values = np.zeros(len(self._dofs), dtype=np.float64)
        self._is_callable = callable(self._value)
        self._refresh()
        self._bc = _DofBC(self._dofs, self._values)

    def _refresh(self):
        if self._is_callable:
            self._values[:] = np.asarray(self._value(self._xT), dtype=np.float64)
        else:
            v = self._value.value if isinstance(self._value, _fem.Constant) else self._value
            self._values[:] = float(v)
        self._version += 1

    def update_bc(self):
        """Re-evaluate a callable value (``bcs.py:128-133``)."""
        if self._is_callable:
            self._refresh()

    def current_values(self) -> np.ndarray:
        """Values ``set_bc`` would read now (a Constant is read live, like ``dirichletbc(Constant)``)."""
        if not self._is_callable:
            v = float(self._value.value if isinstance(self._value, _fem.Constant) else self._value)
            if len(self._values) and self._values[0] != v:
                self._values[:] = v
                self._version += 1
        return self._values

    def apply(self, x):
        """``set_bc(x, [bc])`` (``bcs.py:135-139``): x[dofs] = g[dofs]."""
        arr = x.array if hasattr(x, "array") else x
        arr[self._dofs] = self.current_values()


def _same_topology(a, b) -> bool:
    """The tags belong to the mesh (``bcs.py:223`` compares ``mesh.topology._cpp_object`` with the tags' topology)."""
    unwrap = lambda t: getattr(t, "_cpp_object", t)
    return a is b or unwrap(a) is unwrap(b) or bool(unwrap(a) == unwrap(b))


def _cell_facets(mesh, fdim: int) -> np.ndarray:
    """(n_local_cells, facets per cell) facet ids in reference-cell order (facet i opposite vertex i): the provider's
    ``cell_entities`` or, for a DOLFINx mesh, its cell-to-facet connectivity (owned and ghost cells)."""
    top = mesh.topology
    if hasattr(top, "cell_entities"):
        return top.cell_entities(fdim)
    conn = top.connectivity(top.dim, fdim)
    if conn is None:
        top.create_connectivity(top.dim, fdim)
        conn = top.connectivity(top.dim, fdim)
    return np.asarray(conn.array).reshape(-1, top.dim + 1)


class PressureBC:
    """Natural pressure condition on tagged facets (``bcs.py:142-268``): contributes
    ``int h n_i dv/dx_i ds`` to the tentative-velocity RHS and a homogeneous Dirichlet condition
    on the pressure correction."""

    def __init__(self, value, marker: tuple[MeshTags, int]):
        self._subdomain_data, self._subdomain_id = marker
        self._value = value

    def create_bcs(self, V: _fem.FunctionSpace, Q: _fem.FunctionSpace):
        mesh = V.mesh
        assert _same_topology(mesh.topology, self._subdomain_data.topology)  # bcs.py:223
        tags = self._subdomain_data
        if isinstance(self._subdomain_id, tuple):
            facets = tags.indices[np.isin(tags.values, np.asarray(self._subdomain_id, dtype=np.int32))]
        else:
            facets = tags.find(np.int32(self._subdomain_id))
        self._facets = np.asarray(facets, dtype=np.int32)
        self._V, self._Q = V, Q
        fdim = mesh.topology.dim - 1
        mesh.topology.create_connectivity(fdim, mesh.topology.dim)
        dofs = _fem.locate_dofs_topological(Q, fdim, self._facets)
        self._bc = _DofBC(dofs, np.zeros(len(dofs)))  # bcs.py:245-253
        # (cell, local facet index) of every tagged facet: what the ds-integral kernel iterates over
        cf = _cell_facets(mesh, fdim)
        mask = np.isin(cf, self._facets)
        cells, local = np.nonzero(mask)
        local_cells = getattr(Q, "_local_cells", None)
        if local_cells is not None:
            # multi-rank: keep the tagged facets of the cells this rank holds, owned AND ghost (a ghost cell's facet
            # contributes to dofs owned here; rows of ghost dofs are dropped by the kernel), in local cell numbering
            g2l = np.full(mesh.geometry.dofmap.shape[0], -1, dtype=np.int64)
            g2l[local_cells] = np.arange(len(local_cells))
            keep = g2l[cells] >= 0
            cells, local = g2l[cells[keep]], local[keep]
        self._facet_cells, self._facet_local = cells.astype(np.int32), local.astype(np.int32)
        self._is_callable = callable(self._value)
        self._h = np.zeros(Q.num_dofs)  # nodal values of the boundary pressure in Q
        self._version = 0
        self.update_bc(force=True)

    def update_bc(self, force: bool = False):
        """``bcs.py:255-260``."""
        if self._is_callable:
            self._h[:] = np.asarray(self._value(self._Q.tabulate_dof_coordinates().T), dtype=np.float64)
            self._version += 1
        else:
            # a float or a Constant: the reference keeps the Constant inside the ds-form (bcs.py:233-242), so a
            # changed `.value` changes the next assembled surface term -- read it live, as DirichletBC does
            v = float(self._value.value if isinstance(self._value, _fem.Constant) else self._value)
            if force or (len(self._h) and self._h[0] != v):
                self._h[:] = v
                self._version += 1

    @property
    def bc(self) -> _DofBC:
        return self._bc

    def rhs(self, i: int):
        """Descriptor of ``value * n_i * v.dx(i) * ds`` (``bcs.py:233-242,266-268``): facets + nodal h."""
        assert i < self._V.mesh.geometry.dim
        return ("p_surf", i, self._facets, self._h)
