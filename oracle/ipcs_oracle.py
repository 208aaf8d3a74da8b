"""CPU ORACLE (test infrastructure -- never imported by the product path).

A numpy/scipy restatement of the reference's IPCS fractional step
(``/root/reference/src/oasisx/fracstep.py:277-705``; algebra in SURVEY.md Appendix A).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this file.

PARITY UNPINNED: the reference's arithmetic lives in DOLFINx/FFCx/basix/PETSc/MUMPS
(``fenics-dolfinx>=0.10``, ``/root/reference/pyproject.toml:13``; PETSc/MUMPS versions unpinned,
``.github/workflows/tests.yml:43``), none of which is installable here, and the reference's tests
hold no golden vectors.  This restatement is therefore pinned only by
  * the sympy-exact reference-element values of SURVEY.md Appendix C (``tests/test_tables.py``),
  * the relational checks the reference's own tests use (mat-vec RHS == direct RHS,
    ``test/test_tentative_velocity.py:235``; ``demo/assembly_strategies.py:142``),
  * the analytic Taylor-Green solution (``demo/taylor_green.py:41-53,176-191``).

Derivation is deliberately independent of the product: element tensors come from Gauss-Jacobi
(Duffy) quadrature of numerically tabulated basis functions, NOT from
``oasisx_b200/csrc/ref_tables.h``; linear systems are solved with sparse LU (``splu``), mirroring
the reference demo's ``preonly+lu`` (``demo/taylor_green.py:117-121``).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.special import roots_jacobi

EDGES = {
    2: [(1, 2), (0, 2), (0, 1)],
    3: [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)],
}


# ----------------------------------------------------------------------------------------
# quadrature and basis tabulation on the reference simplex
# ----------------------------------------------------------------------------------------
def simplex_quadrature(d: int, degree: int):
    """Collapsed-coordinate Gauss-Jacobi rule exact for polynomials of total degree <= degree."""
    n = degree // 2 + 1
    if d == 1:
        x0, w0 = roots_jacobi(n, 0, 0)
        return ((x0 + 1) / 2)[:, None], w0 / 2
    if d == 2:
        x0, w0 = roots_jacobi(n, 0, 0)
        x1, w1 = roots_jacobi(n, 1, 0)
        a, b = (x1 + 1) / 2, (x0 + 1) / 2  # a in (0,1) with weight (1-a)
        A, B = np.meshgrid(a, b, indexing="ij")
        W = np.outer(w1 / 4, w0 / 2)
        pts = np.stack([A.ravel(), (B * (1 - A)).ravel()], axis=1)
        return pts, W.ravel()
    x0, w0 = roots_jacobi(n, 0, 0)
    x1, w1 = roots_jacobi(n, 1, 0)
    x2, w2 = roots_jacobi(n, 2, 0)
    a, b, c = (x2 + 1) / 2, (x1 + 1) / 2, (x0 + 1) / 2
    A, B, C = np.meshgrid(a, b, c, indexing="ij")
    W = np.einsum("i,j,k->ijk", w2 / 8, w1 / 4, w0 / 2)
    X = A
    Y = B * (1 - A)
    Z = C * (1 - A) * (1 - B)
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1), W.ravel()


def tabulate(d: int, degree: int, pts: np.ndarray):
    """Lagrange basis values (npts, nd) and reference gradients (npts, nd, d); basix ordering."""
    lam = np.empty((len(pts), d + 1))
    lam[:, 0] = 1 - pts.sum(axis=1)
    lam[:, 1:] = pts
    dlam = np.zeros((d + 1, d))
    dlam[0] = -1
    dlam[1:] = np.eye(d)
    if degree == 1:
        return lam, np.broadcast_to(dlam, (len(pts), d + 1, d)).copy()
    nd = d + 1 + len(EDGES[d])
    phi = np.empty((len(pts), nd))
    dphi = np.empty((len(pts), nd, d))
    for a in range(d + 1):
        phi[:, a] = lam[:, a] * (2 * lam[:, a] - 1)
        dphi[:, a] = (4 * lam[:, a, None] - 1) * dlam[a]
    for e, (a, b) in enumerate(EDGES[d]):
        phi[:, d + 1 + e] = 4 * lam[:, a] * lam[:, b]
        dphi[:, d + 1 + e] = 4 * (lam[:, a, None] * dlam[b] + lam[:, b, None] * dlam[a])
    return phi, dphi


class Geometry:
    """Affine simplex geometry: J, |det J|, J^{-1} per cell."""

    def __init__(self, x: np.ndarray, cells: np.ndarray, d: int):
        X = x[cells][:, :, :d]  # (nc, d+1, d)
        self.J = np.transpose(X[:, 1:] - X[:, :1], (0, 2, 1))  # J[k, delta] = dx_k/dxi_delta
        self.detJ = np.abs(np.linalg.det(self.J))
        self.Kinv = np.linalg.inv(self.J)  # Kinv[delta, k]
        self.X0 = X[:, 0]
        self.d = d

    def physical_points(self, pts):
        return self.X0[:, None, :] + np.einsum("ckd,qd->cqk", self.J, pts)


def _coo(rows, cols, vals, shape):
    nr, nc = rows.shape[1], cols.shape[1]
    R = np.repeat(rows, nc, axis=1).ravel()
    C = np.tile(cols, (1, nr)).ravel()
    return sp.coo_matrix((vals.reshape(-1), (R, C)), shape=shape).tocsr()


class Forms:
    """The fixed menu of cell integrals of ``fracstep.py:289-358``."""

    def __init__(self, x, cells, d, vdofs, qdofs, nV, nQ, deg_v, deg_q=1):
        self.g = Geometry(x, cells, d)
        self.d, self.vdofs, self.qdofs, self.nV, self.nQ = d, vdofs, qdofs, nV, nQ
        self.deg_v, self.deg_q = deg_v, deg_q
        self.pts, self.w = simplex_quadrature(d, 2 * deg_v + max(deg_v - 1, 1) + 1)
        self.phi, self.dphi = tabulate(d, deg_v, self.pts)
        self.psi, self.dpsi = tabulate(d, deg_q, self.pts)

    # physical gradients: grad phi_j = Kinv^T dref phi_j  -> (nc, q, j, k)
    def _gphys(self, dref):
        return np.einsum("cdk,qjd->cqjk", self.g.Kinv, dref)

    def mass_V(self):
        Me = np.einsum("q,qi,qj->ij", self.w, self.phi, self.phi)
        return _coo(self.vdofs, self.vdofs, self.g.detJ[:, None, None] * Me, (self.nV, self.nV))

    def mass_Q(self):
        Me = np.einsum("q,qi,qj->ij", self.w, self.psi, self.psi)
        return _coo(self.qdofs, self.qdofs, self.g.detJ[:, None, None] * Me, (self.nQ, self.nQ))

    def stiffness_V(self):
        gp = self._gphys(self.dphi)
        Ke = np.einsum("c,q,cqik,cqjk->cij", self.g.detJ, self.w, gp, gp)
        return _coo(self.vdofs, self.vdofs, Ke, (self.nV, self.nV))

    def stiffness_Q(self):
        gp = self._gphys(self.dpsi)
        Ke = np.einsum("c,q,cqik,cqjk->cij", self.g.detJ, self.w, gp, gp)
        return _coo(self.qdofs, self.qdofs, Ke, (self.nQ, self.nQ))

    def convection(self, uab):
        """C[i,j] = int (uab . grad phi_j) phi_i   (``fracstep.py:355-358``)."""
        gp = self._gphys(self.dphi)
        uq = np.stack([np.einsum("qa,ca->cq", self.phi, u[self.vdofs]) for u in uab], axis=2)
        adv = np.einsum("cqk,cqjk->cqj", uq, gp)
        Ce = np.einsum("c,q,qi,cqj->cij", self.g.detJ, self.w, self.phi, adv)
        return _coo(self.vdofs, self.vdofs, Ce, (self.nV, self.nV))

    def p_vdxi(self, i):
        """P_i[j,q] = int psi_q d_i phi_j  (``fracstep.py:311-313``)."""
        gp = self._gphys(self.dphi)[..., i]
        Pe = np.einsum("c,q,cqj,qr->cjr", self.g.detJ, self.w, gp, self.psi)
        return _coo(self.vdofs, self.qdofs, Pe, (self.nV, self.nQ))

    def grad_p(self, i):
        """G_i[j,q] = int d_i psi_q phi_j  (``fracstep.py:348-350``)."""
        gp = self._gphys(self.dpsi)[..., i]
        Ge = np.einsum("c,q,qj,cqr->cjr", self.g.detJ, self.w, self.phi, gp)
        return _coo(self.vdofs, self.qdofs, Ge, (self.nV, self.nQ))

    def divu(self, i):
        """D_i[q,j] = int d_i phi_j psi_q  (``fracstep.py:332-334``)."""
        gp = self._gphys(self.dphi)[..., i]
        De = np.einsum("c,q,qr,cqj->crj", self.g.detJ, self.w, self.psi, gp)
        return _coo(self.qdofs, self.vdofs, De, (self.nQ, self.nV))

    def load_V(self, f: float):
        le = np.einsum("q,qj->j", self.w, self.phi)
        b = np.zeros(self.nV)
        np.add.at(b, self.vdofs.ravel(), (f * self.g.detJ[:, None] * le[None, :]).ravel())
        return b

    def load_Q(self):
        le = np.einsum("q,qj->j", self.w, self.psi)
        b = np.zeros(self.nQ)
        np.add.at(b, self.qdofs.ravel(), (self.g.detJ[:, None] * le[None, :]).ravel())
        return b

    def l2_error_sq(self, comps, dofs_per_comp, exact, degree=10, space="V"):
        """sum_k int (u_h,k - exact_k(x))^2  with a degree-`degree` rule (``demo/taylor_green.py:186-207``)."""
        pts, w = simplex_quadrature(self.d, degree)
        tab, _ = tabulate(self.d, self.deg_v if space == "V" else self.deg_q, pts)
        xq = self.g.physical_points(pts)  # (nc, q, d)
        xq3 = np.zeros((3,) + xq.shape[:2])
        xq3[: self.d] = np.moveaxis(xq, 2, 0)
        err = 0.0
        for uh, ex in zip(comps, exact):
            uq = np.einsum("qa,ca->cq", tab, uh[dofs_per_comp])
            e = uq - ex(xq3.reshape(3, -1)).reshape(uq.shape)
            err += float(np.einsum("c,q,cq->", self.g.detJ, w, e * e))
        return err


FACET_VERTS = {
    2: [(1, 2), (0, 2), (0, 1)],
    3: [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)],
}


def pressure_surface_vector(F: "Forms", i: int, facet_cells, facet_local, h_nodal):
    """int_Gamma h n_i dv/dx_i ds  (``bcs.py:233-242``: rhs(i) = value * n_i * v.dx(i) * ds), h given
    nodally in Q, on the exterior facets (cell, local facet index); numerical facet quadrature."""
    d = F.d
    fpts, fw = simplex_quadrature(d - 1, 3)
    fw = fw / fw.sum()  # weights relative to the facet measure
    ref_verts = np.vstack([np.zeros(d), np.eye(d)])
    b = np.zeros(F.nV)
    for c, f in zip(np.asarray(facet_cells), np.asarray(facet_local)):
        fv = ref_verts[list(FACET_VERTS[d][f])]  # (d, d) facet vertices in cell reference coordinates
        mu = np.hstack([1 - fpts.sum(axis=1, keepdims=True), fpts])  # barycentrics on the facet
        pts = mu @ fv
        phi, dphi = tabulate(d, F.deg_v, pts)
        psi, _ = tabulate(d, F.deg_q, pts)
        Kinv, detJ = F.g.Kinv[c], F.g.detJ[c]
        glam = Kinv.T @ np.concatenate([[-np.ones(d)], np.eye(d)])[f]  # physical grad of lambda_f (inward)
        area = detJ * np.linalg.norm(glam) / (2.0 if d == 3 else 1.0)
        n = -glam / np.linalg.norm(glam)
        gphys = np.einsum("dk,qjd->qjk", Kinv, dphi)  # (q, j, k)
        hq = psi @ h_nodal[F.qdofs[c]]
        contrib = area * np.einsum("q,q,qj->j", fw, hq, gphys[:, :, i]) * n[i]
        np.add.at(b, F.vdofs[c], contrib)
    return b


def zero_rows(A: sp.csr_matrix, rows: np.ndarray, diag: float) -> sp.csr_matrix:
    """``Mat.zeroRowsLocal(rows, diag)``: zero the rows, keep the pattern, put `diag` on the diagonal."""
    A = A.tocsr(copy=True)
    for r in rows:
        s, e = A.indptr[r], A.indptr[r + 1]
        A.data[s:e] = 0.0
        A.data[s + np.searchsorted(A.indices[s:e], r)] = diag
    return A


def zero_rows_cols(A: sp.csr_matrix, rows: np.ndarray, diag: float) -> sp.csr_matrix:
    """``assemble_matrix(..., bcs=)`` result: BC rows and columns zeroed, unit diagonal (Appendix D)."""
    mask = np.ones(A.shape[0])
    mask[rows] = 0.0
    Dm = sp.diags(mask)
    B = (Dm @ A @ Dm).tocsr()
    B = B + sp.csr_matrix((np.full(len(rows), diag), (rows, rows)), shape=A.shape)
    return B.tocsr()


class OracleIPCS:
    """One object == one ``FractionalStep_AB_CN`` (``fracstep.py:29``), all state in numpy.

    bcs_u[i] is a list of ``(dofs, value)`` with value a float, an array over all dofs of the
    component space, or a callable ``f(x)`` re-evaluated by :meth:`update_bcs`; bcs_p is a list of
    pressure-BC dof arrays (homogeneous Dirichlet on the pressure correction, ``bcs.py:245-253``)
    optionally with surface vectors ``p_surf[i]`` (``fracstep.py:461-465``)."""

    def __init__(self, x, cells, d, vdofs, qdofs, xV, xQ, deg_v, bcs_u, bcs_p=(), body_force=None,
                 low_memory=False, rotational=False, p_surf=None, pressure_facets=()):
        self.d = d
        self.nV, self.nQ = xV.shape[0], xQ.shape[0]
        self.xV, self.xQ = xV, xQ
        self.vdofs, self.qdofs = vdofs.astype(np.int64), qdofs.astype(np.int64)
        self.F = Forms(x, cells, d, self.vdofs, self.qdofs, self.nV, self.nQ, deg_v)
        self.bcs_u = bcs_u
        self.bcs_p = [np.asarray(b, dtype=np.int64) for b in bcs_p]
        self.p_surf = p_surf
        # natural pressure BCs: list of (facet_cells, facet_local, value) with value float or callable(x)
        self.pressure_facets = list(pressure_facets)
        self.rotational = rotational
        z = lambda n: np.zeros(n)
        self.u = [z(self.nV) for _ in range(d)]
        self.u1 = [z(self.nV) for _ in range(d)]
        self.u2 = [z(self.nV) for _ in range(d)]
        self.ps, self.p, self.dp, self.b2 = z(self.nQ), z(self.nQ), z(self.nQ), z(self.nQ)
        self.rhs1 = [z(self.nV) for _ in range(d)]
        self.b_first = [z(self.nV) for _ in range(d)]
        self.bc_vals = [[None] * len(b) for b in bcs_u]
        self.update_bcs()
        # ---- _preassemble (fracstep.py:360-409)
        F = self.F
        self.M, self.K = F.mass_V(), F.stiffness_V()
        self.Ap_plain = F.stiffness_Q()
        pdofs = np.unique(np.concatenate(self.bcs_p)) if self.bcs_p else np.zeros(0, np.int64)
        self.pdofs = pdofs
        self.Ap = zero_rows_cols(self.Ap_plain, pdofs, 1.0) if len(pdofs) else self.Ap_plain
        body_force = (0.0,) * d if body_force is None else body_force
        self.b0 = [F.load_V(float(f)) for f in body_force]
        self.P = [F.p_vdxi(i) for i in range(d)]
        self.G = [F.grad_p(i) for i in range(d)]
        self.D = [F.divu(i) for i in range(d)]
        self.mQ = F.load_Q()
        self.MQ = F.mass_Q() if rotational else None
        self._lu_M = spla.splu(self.M.tocsc())
        self._lu_Ap = None
        self.A = None

    # ---- boundary values -----------------------------------------------------------------
    def update_bcs(self):
        """``bc.update_bc()`` (``bcs.py:128-133``) restricted to the BC dofs (``set_bc`` reads no others)."""
        for i, bcl in enumerate(self.bcs_u):
            for k, (dofs, val) in enumerate(bcl):
                if callable(val):
                    self.bc_vals[i][k] = np.asarray(val(self.xV[dofs].T), dtype=np.float64)
                elif np.ndim(val) == 0:
                    self.bc_vals[i][k] = np.full(len(dofs), float(val))
                else:
                    self.bc_vals[i][k] = np.asarray(val)[dofs]

    # ---- stages --------------------------------------------------------------------------
    def assemble_first(self, dt, nu):
        if self.pressure_facets:  # :445-446 update, :461-465 surface vector
            self.p_surf = [np.zeros(self.nV) for _ in range(self.d)]
            for fc, fl, val in self.pressure_facets:
                h = np.asarray(val(self.xQ.T), dtype=np.float64) if callable(val) else np.full(self.nQ, float(val))
                for i in range(self.d):
                    self.p_surf[i] += pressure_surface_vector(self.F, i, fc, fl, h)
        uab = [1.5 * a - 0.5 * b for a, b in zip(self.u1, self.u2)]  # :432-434
        C = self.F.convection(uab)  # :435-437
        R = (-0.5) * C + (1.0 / dt) * self.M + (-0.5 * nu) * self.K  # :438-442
        self.R = R
        for i in range(self.d):  # :449-465
            self.b_first[i] = R @ self.u1[i] + self.b0[i]
            if self.p_surf is not None:
                self.b_first[i] = self.b_first[i] + self.p_surf[i]
        A = (-1.0) * R + (2.0 / dt) * self.M  # :468-469
        rows = np.unique(np.concatenate([b[0] for b in self.bcs_u[0]])) if self.bcs_u[0] else []
        self.A = zero_rows(A.tocsr(), np.asarray(rows, dtype=np.int64), 1.0)  # :471-472
        self._lu_A = None

    def velocity_tentative_assemble(self):
        for i in range(self.d):  # :499-506
            self.rhs1[i] = self.b_first[i] + self.P[i] @ self.ps

    def velocity_tentative_solve(self):
        if self._lu_A is None:
            self._lu_A = spla.splu(self.A.tocsc())
        diff = 0.0
        for i in range(self.d):
            for (dofs, _), vals in zip(self.bcs_u[i], self.bc_vals[i]):  # :517-518
                self.rhs1[i][dofs] = vals
            old = self.u[i].copy()
            self.u[i] = self._lu_A.solve(self.rhs1[i])  # :521
            diff += np.linalg.norm(old - self.u[i])  # :523-524
        return diff, np.full(self.d, 4, dtype=np.int32)

    def pressure_assemble(self, dt):
        b2 = np.zeros(self.nQ)
        for i in range(self.d):  # :540-542
            b2 += self.D[i] @ self.u[i]
        b2 *= -1.0 / dt  # :546
        b2[self.pdofs] = 0.0  # :549-550
        self.b2 = b2

    def pressure_solve(self, nu=None):
        if len(self.pdofs) == 0:
            self.b2 = self.b2 - self.b2.mean()  # MatNullSpaceRemove, :573-574
            if self._lu_Ap is None:  # bordered system fixes the constant; removed again below
                one = np.ones((self.nQ, 1))
                aug = sp.bmat([[self.Ap, sp.csr_matrix(one)], [sp.csr_matrix(one.T), None]]).tocsc()
                self._lu_Ap = spla.splu(aug)
            self.dp = self._lu_Ap.solve(np.concatenate([self.b2, [0.0]]))[:-1]
            self.dp = self.dp - (self.mQ @ self.dp) / self.mQ.sum()  # :579-591
        else:
            if self._lu_Ap is None:
                self._lu_Ap = spla.splu(self.Ap.tocsc())
            self.dp = self._lu_Ap.solve(self.b2)
        if self.rotational:  # :593-602, xi = 0.5 (:238)
            rhs = self.MQ @ (self.p + self.dp)
            for i in range(self.d):
                rhs -= 0.5 * nu * (self.D[i] @ self.u[i])
            self.ps = spla.splu(self.MQ.tocsc()).solve(rhs)
        else:
            self.ps = self.p + self.dp  # :604
        return 4

    def velocity_update(self, dt):
        for i in range(self.d):  # :636-656
            b3 = self.M @ self.u[i] - dt * (self.G[i] @ self.dp)
            self.u[i] = self._lu_M.solve(b3)
        return np.full(self.d, 4, dtype=np.int32)

    def solve(self, dt, nu, max_error=1e-12, max_iter=10):
        inner, diff = 0, 1e8
        self.ps = self.p.copy()  # :673
        self.update_bcs()  # :675
        self.assemble_first(dt, nu)
        while inner < max_iter and diff > max_error:  # :677-684
            inner += 1
            self.velocity_tentative_assemble()
            diff, _ = self.velocity_tentative_solve()
            self.pressure_assemble(dt)
            self.pressure_solve(nu)
        self.velocity_update(dt)
        for i in range(self.d):  # :689-693
            self.u2[i] = self.u1[i].copy()
            self.u1[i] = self.u[i].copy()
        self.p = self.ps.copy()
        return diff
