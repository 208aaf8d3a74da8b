"""Exact reference tensors (tools/gen_ref_tables.py -> csrc/ref_tables.h) against (1) the sympy-exact
known answers of SURVEY.md Appendix C and (2) the oracle's independent Gauss-Jacobi tabulation."""
import os

import numpy as np
import pytest

from oracle.ipcs_oracle import simplex_quadrature, tabulate

T = np.load(os.path.join(os.path.dirname(__file__), "..", "oasisx_b200", "_ref_tables.npz"))


def test_appendix_c_known_answers():
    M = np.round(T["D3P2_MV"] * 2520).astype(int)
    assert list(M[0]) == [6, 1, 1, 1, -6, -6, -6, -4, -4, -4]
    assert list(M[4]) == [-6, -6, -4, -4, 32, 16, 16, 16, 16, 8]
    np.testing.assert_allclose(T["D3P2_MV"].sum(), 1 / 6, rtol=1e-15)
    S = T["D3P2_SV"]
    K = (S[0, 0] + S[1, 1] + S[2, 2]) * 30
    np.testing.assert_allclose(K[0], [9, 1, 1, 1, 2, 2, 2, -6, -6, -6], atol=1e-13)
    np.testing.assert_allclose(K[1], [1, 3, 0, 0, 0, -1, -1, 1, 1, -4], atol=1e-13)
    np.testing.assert_allclose(K[4], [2, 0, -1, -1, 16, 4, 4, -8, -8, -8], atol=1e-13)
    np.testing.assert_allclose(K[7], [-6, 1, 1, -4, -8, -8, -8, 24, 4, 4], atol=1e-13)
    np.testing.assert_allclose(K.sum(axis=1), 0, atol=1e-13)
    C = T["D3P2_T"][:, 0].sum(axis=0) * 360  # w = (1,0,0)
    np.testing.assert_allclose(C[0], [-3, -1, 0, 0, 0, -4, -4, 4, 4, 4], atol=1e-12)
    np.testing.assert_allclose(C[4], [4, -4, 0, 0, 0, 16, 16, -16, -16, 0], atol=1e-12)
    np.testing.assert_allclose(C.sum(axis=1), 0, atol=1e-12)
    Px = T["D3P2_PX"][0] * 120
    np.testing.assert_allclose(Px[0], [-3, 1, 1, 1], atol=1e-13)
    np.testing.assert_allclose(Px[5], [4, 4, 4, 8], atol=1e-13)
    np.testing.assert_allclose(Px[9], [4, -4, 0, 0], atol=1e-13)
    Gx = T["D3P2_GX"][0] * 120
    np.testing.assert_allclose(Gx[0], [1, -1, 0, 0], atol=1e-13)
    np.testing.assert_allclose(Gx[4], [-4, 4, 0, 0], atol=1e-13)
    SQ = T["D3P2_SQ"]
    np.testing.assert_allclose((SQ[0, 0] + SQ[1, 1] + SQ[2, 2]) * 6,
                               [[3, -1, -1, -1], [-1, 1, 0, 0], [-1, 0, 1, 0], [-1, 0, 0, 1]], atol=1e-13)
    M2 = np.round(T["D2P2_MV"] * 360).astype(int)
    assert list(M2[0]) == [6, -1, -1, -4, 0, 0]
    assert list(M2[3]) == [-4, 0, 0, 32, 16, 16]


@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("deg", [1, 2])
def test_tables_match_oracle_quadrature(d, deg):
    pts, w = simplex_quadrature(d, 8)
    phi, dphi = tabulate(d, deg, pts)
    psi, dpsi = tabulate(d, 1, pts)
    tag = f"D{d}P{deg}"
    ref = {
        "MV": np.einsum("q,qi,qj->ij", w, phi, phi),
        "SV": np.einsum("q,qia,qjb->abij", w, dphi, dphi),
        "T": np.einsum("q,qa,qjd,qi->adij", w, phi, dphi, phi),
        "PX": np.einsum("q,qr,qjd->djr", w, psi, dphi),
        "GX": np.einsum("q,qrd,qj->djr", w, dpsi, phi),
        "SQ": np.einsum("q,qia,qjb->abij", w, dpsi, dpsi),
        "MQ": np.einsum("q,qi,qj->ij", w, psi, psi),
        "LV": np.einsum("q,qj->j", w, phi),
        "LQ": np.einsum("q,qj->j", w, psi),
    }
    for k, v in ref.items():
        np.testing.assert_allclose(T[f"{tag}_{k}"], v, atol=2e-15, err_msg=f"{tag}_{k}")
