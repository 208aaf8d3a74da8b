"""State export and checkpointing for the IPCS loop.

The reference writes the fields with ``dolfinx.io.VTXWriter`` (ADIOS2, ``demo/taylor_green.py:183-184,211-216``) and
has no restart file; ADIOS2 does not exist here, so

* :func:`write_vtu` writes one ASCII ``.vtu`` (VTK unstructured grid, readable by ParaView) per call with the mesh
  and nodal point data -- P2 fields on quadratic simplices (VTK cell types 22 / 24), P1 fields are interpolated to
  the P2 nodes;
* :func:`save_checkpoint` / :func:`load_checkpoint` store what a restart needs -- ``u1``, ``u2``, ``p`` and the time
  (``fracstep.py:689-693``) -- as one ``.npz`` per rank, keyed by the GLOBAL dof numbers so that a restart may use a
  different number of ranks.
"""
from __future__ import annotations

import numpy as np

__all__ = ["write_vtu", "save_checkpoint", "load_checkpoint"]

# basix/UFC local dof order -> VTK quadratic simplex node order (vertices, then edges 01 12 20 [03 13 23])
_VTK_TRI6 = [0, 1, 2, 5, 3, 4]
_VTK_TET10 = [0, 1, 2, 3, 9, 6, 8, 7, 5, 4]


def write_vtu(path: str, space, point_data: dict[str, np.ndarray]):
    """``space``: a scalar P1 or P2 :class:`oasisx_b200.fem.FunctionSpace`; ``point_data``: name -> (n_dofs,) or
    (n_dofs, k) arrays on that space."""
    x = space.tabulate_dof_coordinates()
    cells = np.asarray(space.dofmap.list)
    d = space.mesh.geometry.dim
    if space.degree == 2:
        perm, ctype = (_VTK_TRI6, 22) if d == 2 else (_VTK_TET10, 24)
        cells = cells[:, perm]
    else:
        ctype = 5 if d == 2 else 10
    nc, npc = cells.shape
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">\n<UnstructuredGrid>\n')
        f.write(f'<Piece NumberOfPoints="{x.shape[0]}" NumberOfCells="{nc}">\n<Points>\n<DataArray type="Float64" NumberOfComponents="3" format="ascii">\n')
        np.savetxt(f, x, fmt="%.17g")
        f.write('</DataArray>\n</Points>\n<Cells>\n<DataArray type="Int32" Name="connectivity" format="ascii">\n')
        np.savetxt(f, cells, fmt="%d")
        f.write('</DataArray>\n<DataArray type="Int32" Name="offsets" format="ascii">\n')
        np.savetxt(f, (np.arange(nc) + 1) * npc, fmt="%d")
        f.write('</DataArray>\n<DataArray type="UInt8" Name="types" format="ascii">\n')
        np.savetxt(f, np.full(nc, ctype), fmt="%d")
        f.write("</DataArray>\n</Cells>\n<PointData>\n")
        for name, v in point_data.items():
            v = np.asarray(v, dtype=np.float64)
            v = v.reshape(x.shape[0], -1)
            if v.shape[1] == 2:
                v = np.hstack([v, np.zeros((v.shape[0], 1))])
            f.write(f'<DataArray type="Float64" Name="{name}" NumberOfComponents="{v.shape[1]}" format="ascii">\n')
            np.savetxt(f, v, fmt="%.17g")
            f.write("</DataArray>\n")
        f.write("</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n")


def save_checkpoint(path: str, solver, t: float):
    """Write ``u1``, ``u2``, ``p`` (owned entries) and the time; multi-rank runs write ``path`` + ``.rank<r>``."""
    lp = solver._lp
    nV, nQ = solver._nV_owned, solver._nQ_owned
    gV = lp.V.l2g[:nV] if lp is not None else np.arange(nV)
    gQ = lp.Q.l2g[:nQ] if lp is not None else np.arange(nQ)
    data = {"t": float(t), "gV": gV, "gQ": gQ, "p": solver._p.x.array_ro()[:nQ].copy()}
    for i in range(len(solver._u1)):
        data[f"u1_{i}"] = solver._u1[i].x.array_ro()[:nV].copy()
        data[f"u2_{i}"] = solver._u2[i].x.array_ro()[:nV].copy()
    np.savez(path if lp is None else f"{path}.rank{solver._rank}", **data)


def load_checkpoint(paths, solver) -> float:
    """Restore the state from one or several checkpoint files (any number of writing ranks); returns the time."""
    if isinstance(paths, str):
        paths = [paths]
    lp = solver._lp
    nVl, nQl = solver._Vi[0][0].num_dofs, solver._Q.num_dofs
    posV = {int(g): i for i, g in enumerate(lp.V.l2g)} if lp is not None else None
    posQ = {int(g): i for i, g in enumerate(lp.Q.l2g)} if lp is not None else None
    t = None
    for path in paths:
        with np.load(path if path.endswith(".npz") else path + ".npz") as z:
            t = float(z["t"])
            for gkey, pos, n, names, funcs in (
                ("gV", posV, nVl, [f"u1_{i}" for i in range(len(solver._u1))] + [f"u2_{i}" for i in range(len(solver._u2))],
                 list(solver._u1) + list(solver._u2)),
                ("gQ", posQ, nQl, ["p"], [solver._p]),
            ):
                g = z[gkey]
                if pos is None:
                    idx, sel = g, np.ones(len(g), bool)
                else:
                    idx = np.array([pos.get(int(k), -1) for k in g])
                    sel = idx >= 0
                for name, fn in zip(names, funcs):
                    fn.x.array[idx[sel]] = z[name][sel]
    return t
