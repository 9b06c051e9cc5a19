"""The eigenvalue algorithm of the DI kernel (csrc/score.cu di_eig_kernel: warp-level Householder tridiagonalisation + lane-level
implicit QL), pinned on the CPU through its line-by-line Python model against numpy.linalg.eigvalsh and against the oracle's DI."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import di_eig_model as model  # noqa: E402


@pytest.mark.parametrize("s", [1, 2, 3, 4, 19, 20, 24, 30])
def test_tridiagonal_ql_matches_eigvalsh(s):
    rng = np.random.default_rng(s)
    for trial in range(12):
        G = rng.standard_normal((s, s)) * rng.uniform(1e-3, 3.0)
        if trial % 4 == 1:
            G[:, : s // 2] = 0.0                      # rank deficient
        if trial % 4 == 2:
            G = np.diag(rng.standard_normal(s))       # already diagonal: every reflector is skipped
        if trial % 4 == 3:
            G *= 1e-150                               # squares underflow: rows are treated as zero, eigenvalues < 1e-290
        V = G.T @ G
        got = np.sort(model.ql_eigenvalues(*model.tridiagonalise(V)))
        ref = np.linalg.eigvalsh(V)
        scale = max(float(np.abs(ref).max()), 1e-280)
        assert np.abs(got - ref).max() <= 1e-14 * scale + 1e-290


def test_tridiagonal_form_is_similar():
    rng = np.random.default_rng(7)
    G = rng.standard_normal((20, 20))
    V = G.T @ G
    d, e = model.tridiagonalise(V)
    T = np.diag(d) + np.diag(e[:-1], 1) + np.diag(e[:-1], -1)
    assert np.abs(np.linalg.eigvalsh(T) - np.linalg.eigvalsh(V)).max() <= 1e-13 * np.abs(V).max()
    assert e[-1] == 0.0


def test_di_of_blocks_matches_the_oracle(orc):
    q, s, L = 21, 20, 6
    n = s * L
    rng = np.random.default_rng(3)
    A = rng.standard_normal((n, 3 * n))
    C = A @ A.T / (3 * n) + 0.1 * np.eye(n)
    mJ = orc.inv_cholesky(C)
    ref = orc.compute_DI_gauss(mJ, C, q)
    Lc = [np.linalg.cholesky(C[i * s:(i + 1) * s, i * s:(i + 1) * s]) for i in range(L)]
    for i in range(L - 1):
        for j in range(i + 1, L):
            di = model.di_from_block(mJ[i * s:(i + 1) * s, j * s:(j + 1) * s], Lc[i], Lc[j])
            assert abs(di - ref[i, j]) <= 1e-11 * max(1.0, abs(ref[i, j]))
    assert abs(model.di_from_block(np.zeros((s, s)), Lc[0], Lc[1])) <= 1e-13   # zero coupling: s/2 log(1/2) + s/2 log 2
