"""Slice schedules of the brick SpMM experiment (were in oasisx_b200/fem.py)."""
import numpy as np


def brick_schedule(x: np.ndarray, n_rows: int, lattice=None, tile=(2, 2), rows_per_slice: int = 32):
    """Schedule and grouping of the 32-row slices for the BRICK form of a sliced-ELL operator (csrc/bricks.hpp):
    slices sorted by the small spatial box of their first row -- one slice length in x, `tile` cells in y and z -- and
    the offsets of the groups of equal box (`hint_ptr`).  The library packs the slices of one group into bricks whose
    distinct columns fit a shared-memory gather list.  On a P2 lattice one box holds 2 x 4 x 4 x-lines of dofs = 32
    slices whose rows touch about 4400 distinct columns: 4-5 loads of x per row instead of 28.5 gathers.  Without
    lattice information the box edge is estimated from the bounding box and the dof count (as `slice_order` does).
    Returns (order, hint_ptr)."""
    n_slices = (n_rows + rows_per_slice - 1) // rows_per_slice
    first = x[np.arange(n_slices) * rows_per_slice]
    if lattice is not None:
        p0, h = lattice
    else:
        p0 = x.min(axis=0)
        ext = np.maximum(x.max(axis=0) - p0, 1e-300)
        dim = int(np.count_nonzero(ext > 1e-12 * ext.max()))
        h = np.where(ext > 1e-12 * ext.max(), ext / max(n_rows ** (1.0 / max(dim, 1)) / 2.0, 1.0), 1.0)
    q = np.floor((first - p0) / h + 1e-9).astype(np.int64)
    kz, ky, kx = q[:, 2] // tile[1], q[:, 1] // tile[0], q[:, 0] // rows_per_slice
    order = np.lexsort((np.arange(n_slices), kx, ky, kz))
    kz, ky, kx = kz[order], ky[order], kx[order]
    new = np.ones(n_slices, dtype=bool)
    new[1:] = (kz[1:] != kz[:-1]) | (ky[1:] != ky[:-1]) | (kx[1:] != kx[:-1])
    hint_ptr = np.concatenate([np.flatnonzero(new), [n_slices]])
    return order.astype(np.int32), hint_ptr.astype(np.int32)


def pad_order(order: np.ndarray, hint_ptr: np.ndarray, group: int = 32) -> np.ndarray:
    """The schedule `order` with every hint group padded by -1 entries to a multiple of `group`: a thread block of
    `group` warps then always works on the slices of ONE spatial box (csrc/linalg.cuh: k_spmm)."""
    sizes = np.diff(hint_ptr)
    padded = (sizes + group - 1) // group * group
    out = np.full(int(padded.sum()), -1, dtype=np.int32)
    start = np.concatenate([[0], np.cumsum(padded)[:-1]])
    pos = np.repeat(start - hint_ptr[:-1], sizes) + np.arange(len(order))
    out[pos] = order
    return out


