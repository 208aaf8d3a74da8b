"""Brick form of the sliced-ELL operators (oasisx_b200/csrc/bricks.hpp, b2_host_build_bricks): the host builder checked on
the CPU by executing, in numpy, exactly the indexing k_spmm_brick uses (gather list -> shared memory -> 16-bit
positions) and comparing with the CSR product."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from oasisx_b200 import _lib, fem, mesh as bmesh


def sell_from_csr(indptr, indices, n_rows, n_cols):
    """SELL-32 slice offsets and padded columns as csrc/linalg.cuh: k_sell_slice_len / k_sell_fill_cols build them."""
    ns = (n_rows + 31) // 32
    rl = np.diff(indptr)
    slice_ptr = np.zeros(ns + 1, np.int32)
    for s in range(ns):
        slice_ptr[s + 1] = slice_ptr[s] + 32 * int(rl[s * 32:min(n_rows, s * 32 + 32)].max())
    scols = np.zeros(slice_ptr[-1], np.int32)
    slot_of = np.empty(len(indices), np.int64)  # CSR position -> slot
    for r in range(n_rows):
        s = r >> 5
        length = (slice_ptr[s + 1] - slice_ptr[s]) >> 5
        n = rl[r]
        base = slice_ptr[s] + (r & 31)
        scols[base + 32 * np.arange(length)] = r if r < n_cols else 0
        scols[base + 32 * np.arange(n)] = indices[indptr[r]:indptr[r + 1]]
        slot_of[indptr[r]:indptr[r + 1]] = base + 32 * np.arange(n)
    return slice_ptr, scols, slot_of


def host_build(lib, n_rows, n_cols, slice_ptr, scols, order, hint_ptr, cap, max_slices, threads):
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    nb, ng = np.zeros(1, np.int64), np.zeros(1, np.int64)
    args = [n_rows, n_cols, p(slice_ptr), p(scols), p(order), len(hint_ptr) - 1, p(hint_ptr), cap, max_slices, threads]
    rc = lib.b2_host_build_bricks(*args, p(nb), p(ng), None, None, None, None, 0, 0, None, None)
    if rc != 0:
        return rc, None
    bptr, gptr = np.zeros(nb[0] + 1, np.int32), np.zeros(nb[0] + 1, np.int32)
    glist, lcols = np.zeros(ng[0], np.int32), np.zeros(len(scols), np.uint16)
    grid = min(GRID, int(nb[0]))
    wdesc, wseq = np.zeros(((n_rows + 31) // 32, 4), np.int32), np.zeros(grid * WARPS + 1, np.int32)
    rc = lib.b2_host_build_bricks(*args, p(nb), p(ng), p(bptr), p(gptr), p(glist), p(lcols), WARPS, GRID, p(wdesc), p(wseq))
    host_build.work_lists = (wdesc, wseq, grid)
    return rc, (bptr, gptr, glist, lcols)


WARPS, GRID = 16, 5


def brick_product(n_rows, slice_ptr, vals, order, bptr, gptr, glist, lcols, x, cap):
    """What k_spmm_brick computes, index for index (one component)."""
    y = np.full(n_rows, np.nan)
    for b in range(len(bptr) - 1):
        g = glist[gptr[b]:gptr[b + 1]]
        assert len(g) <= cap
        sx = x[g]  # the fill phase
        for j in range(bptr[b], bptr[b + 1]):
            s = order[j]
            base, length = slice_ptr[s], (slice_ptr[s + 1] - slice_ptr[s]) >> 5
            for lane in range(32):
                row = (s << 5) + lane
                if row >= n_rows:
                    continue
                acc = 0.0
                for t in range(length):
                    slot = base + (t << 5) + lane
                    acc += vals[slot] * sx[lcols[slot] // 8]  # stored as byte offsets
                y[row] = acc
    return y


@pytest.mark.parametrize("case", ["box_lattice", "box_generic", "square", "tiny_cap"])
def test_host_brick_builder_reproduces_the_csr_product(lib, case):
    rng = np.random.default_rng(3)
    if case == "square":
        msh = bmesh.create_rectangle(None, [[0, 0], [1, 2]], [40, 9])
    else:
        msh = bmesh.create_box(None, [[0, 0, 0], [1, 1, 2]], [35, 3, 4])
    if case == "box_generic":
        msh._dof_order = "generic"
    V = fem.functionspace(msh, ("Lagrange", 2))
    n = V.num_dofs
    indptr, indices = fem.build_csr_pattern(V.dofmap.list, V.dofmap.list, n, n)
    slice_ptr, scols, slot_of = sell_from_csr(indptr, indices, n, n)
    lat = msh._lattice if case != "box_generic" else None
    order, hints = fem.brick_schedule(V.tabulate_dof_coordinates(), n, lat)
    ns = (n + 31) // 32
    assert sorted(order.tolist()) == list(range(ns)) and hints[0] == 0 and hints[-1] == ns and np.all(np.diff(hints) > 0)
    cap, max_slices = (4352, 32) if case != "tiny_cap" else (1400, 5)
    rc, out = host_build(lib, n, n, slice_ptr, scols, order, hints, cap, max_slices, 3)
    assert rc == 0
    bptr, gptr, glist, lcols = out
    # bricks tile the schedule, respect the hints, the slice limit and the capacity; gather lists sorted and distinct
    assert bptr[0] == 0 and bptr[-1] == ns and np.all(np.diff(bptr) > 0) and np.diff(bptr).max() <= max_slices
    assert np.isin(hints, bptr).all()
    assert np.diff(gptr).max() <= cap
    for b in range(len(bptr) - 1):
        g = glist[gptr[b]:gptr[b + 1]]
        assert np.all(np.diff(g) > 0)
    # every slot points at its own column
    brick_of_slice = np.empty(ns, np.int64)
    for b in range(len(bptr) - 1):
        brick_of_slice[order[bptr[b]:bptr[b + 1]]] = b
    slot_slice = np.repeat(np.arange(ns), np.diff(slice_ptr))
    assert np.all(lcols % 8 == 0)
    np.testing.assert_array_equal(glist[gptr[brick_of_slice[slot_slice]] + lcols.astype(np.int64) // 8], scols)
    # and the product, executed the way the kernel indexes it, is the CSR product bit for bit
    a = rng.standard_normal(len(indices))
    vals = np.zeros(len(scols))
    vals[slot_of] = a
    x = rng.standard_normal(n)
    y = brick_product(n, slice_ptr, vals, order, bptr, gptr, glist, lcols, x, cap)
    ref = np.array([np.sum(np.cumsum(a[indptr[r]:indptr[r + 1]] * x[indices[indptr[r]:indptr[r + 1]]])[-1:]) for r in range(n)])
    np.testing.assert_allclose(y, sp.csr_matrix((a, indices, indptr), shape=(n, n)) @ x, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(y, ref, rtol=1e-13, atol=1e-13)
    # work lists of the pipelined kernel: block g owns bricks g, g + grid, ...; the slices of a brick are dealt to the
    # warps longest first; a warp's descriptors {brick, slice, first slot, steps} come in the order it meets them
    wdesc, wseq, grid = host_build.work_lists
    assert wseq[0] == 0 and wseq[-1] == ns and np.all(np.diff(wseq) >= 0)
    slen = np.diff(slice_ptr) // 32
    assert np.array_equal(wdesc[:, 2], slice_ptr[wdesc[:, 1]]) and np.array_equal(wdesc[:, 3], slen[wdesc[:, 1]])
    assert sorted(wdesc[:, 1].tolist()) == list(range(ns))
    brick_of = np.empty(ns, np.int64)
    for b in range(len(bptr) - 1):
        brick_of[order[bptr[b]:bptr[b + 1]]] = b
    assert np.array_equal(wdesc[:, 0], brick_of[wdesc[:, 1]])
    loads = np.zeros((len(bptr) - 1, WARPS), np.int64)
    for g in range(grid):
        for w in range(WARPS):
            mine = wdesc[wseq[g * WARPS + w]:wseq[g * WARPS + w + 1]]
            assert np.all(mine[:, 0] % grid == g) and np.all(np.diff(mine[:, 0]) >= 0)  # its block's bricks, in order
            np.add.at(loads[:, w], mine[:, 0], mine[:, 3])
    for b in range(len(bptr) - 1):  # LPT: no warp is ahead of another by more than one slice
        assert loads[b].max() - loads[b].min() <= slen[order[bptr[b]:bptr[b + 1]]].max()
    # one thread or several: the same bricks
    rc1, out1 = host_build(lib, n, n, slice_ptr, scols, order, hints, cap, max_slices, 1)
    assert rc1 == 0 and all(np.array_equal(u, v) for u, v in zip(out, out1))


def test_host_brick_builder_rejects_a_slice_that_cannot_fit(lib):
    msh = bmesh.create_box(None, [[0, 0, 0], [1, 1, 1]], [4, 4, 4])
    V = fem.functionspace(msh, ("Lagrange", 2))
    n = V.num_dofs
    indptr, indices = fem.build_csr_pattern(V.dofmap.list, V.dofmap.list, n, n)
    slice_ptr, scols, _ = sell_from_csr(indptr, indices, n, n)
    order, hints = fem.brick_schedule(V.tabulate_dof_coordinates(), n, msh._lattice)
    rc, _ = host_build(lib, n, n, slice_ptr, scols, order, hints, 64, 32, 2)
    assert rc == -3


def test_brick_gather_ratio_on_a_lattice():
    """The point of the format: on the P2 lattice a row needs ~5 staged vector entries instead of ~28 gathers."""
    msh = bmesh.create_box(None, [[0, 0, 0], [1, 1, 1]], [32, 8, 8])
    V = fem.functionspace(msh, ("Lagrange", 2))
    n = V.num_dofs
    indptr, indices = fem.build_csr_pattern(V.dofmap.list, V.dofmap.list, n, n)
    slice_ptr, scols, _ = sell_from_csr(indptr, indices, n, n)
    order, hints = fem.brick_schedule(V.tabulate_dof_coordinates(), n, msh._lattice)
    rc, (bptr, gptr, glist, lcols) = host_build(_lib.load_library(), n, n, slice_ptr, scols, order, hints, 4352, 32, 4)
    assert rc == 0
    assert len(glist) / n < 7.5 < len(indices) / n / 3
