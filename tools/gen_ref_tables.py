#!/usr/bin/env python3
"""Generate exact reference-element tensors for Lagrange P1/P2 on triangles and tetrahedra.

All integrals are evaluated in exact rational arithmetic (``fractions.Fraction``) with the
Dirichlet formula  int_ref x^a y^b z^c = a! b! c! / (a+b+c+d)!  and are written out as

* ``oasisx_b200/csrc/ref_tables.h``  – ``__constant__``-ready C arrays used by the CUDA kernels
  (and by the C++ CPU port under ``oracle/``), and
* ``oasisx_b200/_ref_tables.npz``    – the same numbers for the Python host layer.

These replace the FFCx-generated ``tabulate_tensor`` kernels the reference JIT-compiles at
``/root/reference/src/oasisx/fracstep.py:289-358`` (forms: mass, stiffness, p*dv/dx_i,
grad p . grad q, du/dx_i*q, dp/dx_i*v, convection).  Local dof ordering follows basix/UFC
(SURVEY.md Appendix C): vertices first, then edges e0=(2,3), e1=(1,3), e2=(1,2), e3=(0,3),
e4=(0,2), e5=(0,1) on the tetrahedron and e0=(1,2), e1=(0,2), e2=(0,1) on the triangle.

The numpy oracle (``oracle/ipcs_oracle.py``) does NOT read these tables: it integrates with
Gauss-Jacobi quadrature, so the two derivations check each other in ``tests/test_tables.py``.
"""
from __future__ import annotations

import itertools
import os
import sys
from fractions import Fraction
from math import factorial

import numpy as np


class Poly:
    """Multivariate polynomial with Fraction coefficients; monomials keyed by exponent tuples."""

    __slots__ = ("d", "c")

    def __init__(self, d, c=None):
        self.d = d
        self.c = {k: v for k, v in (c or {}).items() if v != 0}

    @staticmethod
    def const(d, v):
        return Poly(d, {(0,) * d: Fraction(v)})

    @staticmethod
    def var(d, i):
        e = [0] * d
        e[i] = 1
        return Poly(d, {tuple(e): Fraction(1)})

    def __add__(self, o):
        if not isinstance(o, Poly):
            o = Poly.const(self.d, o)
        c = dict(self.c)
        for k, v in o.c.items():
            c[k] = c.get(k, 0) + v
        return Poly(self.d, c)

    __radd__ = __add__

    def __neg__(self):
        return Poly(self.d, {k: -v for k, v in self.c.items()})

    def __sub__(self, o):
        return self + (-o if isinstance(o, Poly) else Poly.const(self.d, -Fraction(o)))

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, o):
        if not isinstance(o, Poly):
            return Poly(self.d, {k: v * Fraction(o) for k, v in self.c.items()})
        c = {}
        for (k1, v1), (k2, v2) in itertools.product(self.c.items(), o.c.items()):
            k = tuple(a + b for a, b in zip(k1, k2))
            c[k] = c.get(k, 0) + v1 * v2
        return Poly(self.d, c)

    __rmul__ = __mul__

    def diff(self, i):
        c = {}
        for k, v in self.c.items():
            if k[i] > 0:
                e = list(k)
                e[i] -= 1
                c[tuple(e)] = c.get(tuple(e), 0) + v * k[i]
        return Poly(self.d, c)

    def integrate(self):
        """Integral over the reference simplex {x_i >= 0, sum x_i <= 1}."""
        s = Fraction(0)
        for k, v in self.c.items():
            num = 1
            for e in k:
                num *= factorial(e)
            s += v * Fraction(num, factorial(sum(k) + self.d))
        return s


def barycentrics(d):
    lam = [Poly.const(d, 1)]
    for i in range(d):
        lam[0] = lam[0] - Poly.var(d, i)
        lam.append(Poly.var(d, i))
    return lam


EDGES = {
    2: [(1, 2), (0, 2), (0, 1)],
    3: [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)],
}


def lagrange_basis(d, degree):
    lam = barycentrics(d)
    if degree == 1:
        return lam
    if degree == 2:
        phi = [l * (2 * l - 1) for l in lam]
        phi += [4 * lam[a] * lam[b] for (a, b) in EDGES[d]]
        return phi
    raise ValueError(degree)


def tensor(shape, fn):
    out = np.empty(shape, dtype=object)
    for idx in itertools.product(*[range(s) for s in shape]):
        out[idx] = fn(*idx)
    return out


def tables(d, deg_v, deg_q=1):
    phi = lagrange_basis(d, deg_v)
    psi = lagrange_basis(d, deg_q)
    nv, nq = len(phi), len(psi)
    dphi = [[p.diff(k) for k in range(d)] for p in phi]
    dpsi = [[p.diff(k) for k in range(d)] for p in psi]
    t = {}
    t["MV"] = tensor((nv, nv), lambda i, j: (phi[i] * phi[j]).integrate())
    t["SV"] = tensor((d, d, nv, nv), lambda a, b, i, j: (dphi[i][a] * dphi[j][b]).integrate())
    # T[a, delta, i, j] = int phi_a * d_delta phi_j * phi_i
    pp = [[phi[a] * phi[i] for i in range(nv)] for a in range(nv)]
    t["T"] = tensor((nv, d, nv, nv), lambda a, dl, i, j: (pp[a][i] * dphi[j][dl]).integrate())
    # PX[delta, j, q] = int psi_q d_delta phi_j ; GX[delta, j, q] = int d_delta psi_q phi_j
    t["PX"] = tensor((d, nv, nq), lambda dl, j, q: (psi[q] * dphi[j][dl]).integrate())
    t["GX"] = tensor((d, nv, nq), lambda dl, j, q: (dpsi[q][dl] * phi[j]).integrate())
    t["SQ"] = tensor((d, d, nq, nq), lambda a, b, q, r: (dpsi[q][a] * dpsi[r][b]).integrate())
    t["MQ"] = tensor((nq, nq), lambda q, r: (psi[q] * psi[r]).integrate())
    t["LV"] = tensor((nv,), lambda j: phi[j].integrate())
    t["LQ"] = tensor((nq,), lambda q: psi[q].integrate())
    return t


def to_float(arr):
    return np.array([float(x) for x in arr.ravel()], dtype=np.float64).reshape(arr.shape)


def c_array(name, arr):
    flat = to_float(arr).ravel()
    body = ",\n    ".join(
        ", ".join(repr(float(v)) for v in flat[i : i + 6]) for i in range(0, len(flat), 6)
    )
    dims = "".join(f"[{s}]" for s in arr.shape)
    return f"B2_TABLE_QUAL double {name}{dims} = {{\n    {body}}};\n"


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    npz = {}
    hdr = [
        "// GENERATED by tools/gen_ref_tables.py -- do not edit.\n",
        "// Exact reference-simplex tensors for Lagrange P1/P2 (basix/UFC local ordering).\n",
        "// Replaces the FFCx tabulate_tensor kernels behind\n",
        "// /root/reference/src/oasisx/fracstep.py:289-358.\n",
        "#pragma once\n#ifndef B2_TABLE_QUAL\n#define B2_TABLE_QUAL static const\n#endif\n\n",
    ]
    for d in (2, 3):
        for deg in (1, 2):
            tag = f"D{d}P{deg}"
            t = tables(d, deg)
            hdr.append(f"// ---- dim {d}, velocity degree {deg}, pressure degree 1 ----\n")
            for k, v in t.items():
                npz[f"{tag}_{k}"] = to_float(v)
                hdr.append(c_array(f"REF_{tag}_{k}", v))
            hdr.append("\n")
            nnzT = sum(1 for x in t["T"].ravel() if x != 0)
            print(f"{tag}: nv={t['MV'].shape[0]} T nnz {nnzT}/{t['T'].size}", file=sys.stderr)
    with open(os.path.join(root, "oasisx_b200", "csrc", "ref_tables.h"), "w") as f:
        f.writelines(hdr)
    np.savez(os.path.join(root, "oasisx_b200", "_ref_tables.npz"), **npz)


if __name__ == "__main__":
    main()
