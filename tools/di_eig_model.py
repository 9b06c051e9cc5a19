"""Line-by-line Python model of the eigenvalue path of csrc/score.cu di_eig_kernel (the DI score, reference call site
src/GaussDCA.jl:37): Householder tridiagonalisation of the symmetric V as the WARP does it (lane k owns element k of the reflector,
row k of p = V u / H and column k of the rank-2 update; no row rescaling) followed by the implicit QL iteration as ONE LANE does
it on (d, e).  Pure numpy / math, used by tests/test_di_eig_model_cpu.py to pin the restated EISPACK tred1 / tql1 pair against
numpy.linalg.eigvalsh without a GPU.  Not part of the product path."""
import math

import numpy as np

EPS = 2.220446049250313e-16


def tridiagonalise(V):
    """-> (d, e): diagonal and sub-diagonal (e[k] couples k and k+1; e[s-1] = 0), the kernel's warp phase."""
    s = V.shape[0]
    V = np.array(V, dtype=np.float64)
    e = np.zeros(s)
    for r in range(s - 1, 0, -1):
        l = r - 1
        if l == 0:
            e[r - 1] = V[r, 0]
            continue
        x = np.zeros(32)
        x[: l + 1] = V[r, : l + 1]
        h = float((x * x).sum())
        if not h >= 1e-290:
            e[r - 1] = 0.0
            continue
        f = x[l]
        g = -math.sqrt(h) if f >= 0.0 else math.sqrt(h)
        e[r - 1] = g
        h -= f * g
        rh = 1.0 / h
        x[l] = f - g
        p = np.zeros(32)
        p[: l + 1] = (V[: l + 1, : l + 1] @ x[: l + 1]) * rh
        K = float((p * x).sum()) * (0.5 * rh)
        q = p - K * x
        V[: l + 1, : l + 1] -= np.outer(x[: l + 1], q[: l + 1]) + np.outer(q[: l + 1], x[: l + 1])
    return np.diag(V).copy(), e


def ql_eigenvalues(d, e):
    """Implicit QL with Wilkinson shift on (d, e), eigenvalues only: the kernel's lane phase."""
    D, E = np.array(d, dtype=np.float64), np.array(e, dtype=np.float64)
    s = len(D)
    for l in range(s):
        for _ in range(60):
            m = l
            while m < s - 1:
                if abs(E[m]) <= EPS * (abs(D[m]) + abs(D[m + 1])):
                    break
                m += 1
            if m == l:
                break
            el, dl = E[l], D[l]
            g = (D[l + 1] - dl) / (2.0 * el)
            r = math.sqrt(g * g + 1.0)
            g = D[m] - dl + el / (g + math.copysign(r, g))
            sn = cs = 1.0
            p = 0.0
            k = m - 1
            while k >= l:
                ek = E[k]
                f, b = sn * ek, cs * ek
                r = math.sqrt(f * f + g * g)
                E[k + 1] = r
                if r == 0.0:
                    D[k + 1] -= p
                    E[m] = 0.0
                    break
                rinv = 1.0 / r
                sn, cs = f * rinv, g * rinv
                g = D[k + 1] - p
                r = (D[k] - g) * sn + 2.0 * cs * b
                p = sn * r
                D[k + 1] = g + p
                g = cs * r - b
                k -= 1
            if r == 0.0 and k >= l:
                continue
            D[l] -= p
            E[l] = g
            E[m] = 0.0
    return D


def di_from_block(B, Lc_i, Lc_j):
    """DI of one site pair as the kernel evaluates it: G = Lc_i' B Lc_j, V = G'G, eigenvalues, log sum."""
    G = Lc_i.T @ B @ Lc_j
    lam = ql_eigenvalues(*tridiagonalise(G.T @ G))
    s = B.shape[0]
    return 0.5 * s * math.log(0.5) + 0.5 * sum(math.log(1.0 + math.sqrt(1.0 + 4.0 * max(x, 0.0))) for x in lam)
