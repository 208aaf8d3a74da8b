"""Quadrature on the reference simplex for the host side of functionals (``assemble_scalar``).

Collapsed-coordinate (Duffy) Gauss-Jacobi product rules, exact for total degree ``degree``; nodes from
the Golub-Welsch eigenvalue problem (numpy only)."""
from __future__ import annotations

import numpy as np


def _gauss_jacobi(n: int, alpha: int):
    """n-point Gauss rule for the weight (1-x)^alpha on [-1, 1] (beta = 0)."""
    k = np.arange(n, dtype=np.float64)
    a, b = float(alpha), 0.0
    if n == 1:
        x = np.array([(b - a) / (a + b + 2)])
    else:
        diag = (b * b - a * a) / ((2 * k + a + b) * (2 * k + a + b + 2) + (k == 0) * (a + b == 0))
        if a + b == 0:
            diag[0] = (b - a) / (a + b + 2)
        kk = k[1:]
        off = 2.0 / (2 * kk + a + b) * np.sqrt(kk * (kk + a) * (kk + b) * (kk + a + b) / ((2 * kk + a + b - 1) * (2 * kk + a + b + 1)))
        J = np.diag(diag) + np.diag(off, 1) + np.diag(off, -1)
        x = np.linalg.eigvalsh(J)
    # weights from the exactness conditions on 1, x, ..., x^(n-1) (small, well conditioned for n <= 10)
    from math import comb

    V = np.vander(x, n, increasing=True).T
    # moments of x^m against (1-x)^alpha on [-1,1] by substituting t = 1 - x
    mom = np.empty(n)
    for m in range(n):
        mom[m] = sum(comb(m, j) * (-1.0) ** j * 2.0 ** (j + alpha + 1) / (j + alpha + 1) for j in range(m + 1))
    w = np.linalg.solve(V, mom)
    return x, w


def simplex_rule(d: int, degree: int):
    """(points (nq, d), weights (nq,)) on the reference simplex, weights summing to 1/d!."""
    n = degree // 2 + 1
    x0, w0 = _gauss_jacobi(n, 0)
    if d == 1:
        return ((x0 + 1) / 2)[:, None], w0 / 2
    x1, w1 = _gauss_jacobi(n, 1)
    if d == 2:
        a, b = (x1 + 1) / 2, (x0 + 1) / 2
        A, B = np.meshgrid(a, b, indexing="ij")
        return np.stack([A.ravel(), (B * (1 - A)).ravel()], axis=1), np.outer(w1 / 4, w0 / 2).ravel()
    x2, w2 = _gauss_jacobi(n, 2)
    a, b, c = (x2 + 1) / 2, (x1 + 1) / 2, (x0 + 1) / 2
    A, B, Cc = np.meshgrid(a, b, c, indexing="ij")
    W = np.einsum("i,j,k->ijk", w2 / 8, w1 / 4, w0 / 2)
    return np.stack([A.ravel(), (B * (1 - A)).ravel(), (Cc * (1 - A) * (1 - B)).ravel()], axis=1), W.ravel()
