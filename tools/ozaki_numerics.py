"""CPU study for the next step of the inversion: how many INT8 slices does an Ozaki-style emulation of
the big FP64 GEMMs of the blocked Cholesky + inverse need so that mJ = inv(C) stays within the 1e-9 normwise tolerance?

The blocked algorithm of csrc/chol.cu is replayed in numpy (diagonal blocks and panel products in FP64, exactly as the GPU
keeps them on DMMA; trailing updates, the trtri levels and the lauum product through `sliced_gemm`), on covariance matrices
produced by the CPU oracle from the synthetic generator of SURVEY 8(d).

sliced_gemm(A, B) ~ A @ B.T:  every row of A and of B is scaled by a power of two to [-1, 1) and cut into `s` signed digits of
`w` bits (int8 for w <= 7); digit products A_t @ B_u.T are exact integers (what tcgen05 kind::i8 with S32 accumulation
delivers); only pairs t + u < s are formed (s (s + 1) / 2 products); the result is summed in FP64.

    python tools/ozaki_numerics.py [L] [M]
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as g  # noqa: E402

NB = 128


def slices(A, w, s):
    """rows of A -> exponents e (A = 2^e * A'), and s digit matrices D_t with A' ~ sum_t D_t 2^(-w (t+1)), |D_t| <= 2^(w-1)."""
    amax = np.max(np.abs(A), axis=1)
    e = (np.frexp(amax)[1] + 1).astype(np.float64)          # amax = f 2^e0, f in [1/2, 1)  ->  |A / 2^(e0+1)| < 1/2, exactly
    R = A / np.exp2(e)[:, None]
    D = []
    for t in range(s):
        scale = np.exp2(w * (t + 1))
        d = np.rint(R * scale)
        R = R - d / scale
        D.append(d)
    return e, D


def sliced_gemm(A, B, w, s):
    ea, DA = slices(A, w, s)
    eb, DB = slices(B, w, s)
    acc = np.zeros((A.shape[0], B.shape[0]))
    for t in range(s):
        for u in range(s - t):
            acc += (DA[t] @ DB[u].T) * np.exp2(-w * (t + u + 2))      # exact integers times a power of two
    return acc * np.exp2(ea)[:, None] * np.exp2(eb)[None, :]


def blocked_inverse(C, gemm):
    """csrc/chol.cu in numpy: potrf (NB = 128, right-looking), X = inv(L) by recursive doubling, mJ = X' X."""
    n = C.shape[0]
    nb = (n + NB - 1) // NB
    npad = nb * NB
    A = np.eye(npad)
    A[:n, :n] = C
    X = np.zeros((npad, npad))
    blk = lambda i: slice(i * NB, (i + 1) * NB)  # noqa: E731
    for k in range(nb):
        Lkk = np.linalg.cholesky(A[blk(k), blk(k)])
        A[blk(k), blk(k)] = Lkk
        X[blk(k), blk(k)] = np.linalg.inv(Lkk)
        if k + 1 < nb:
            lo = slice((k + 1) * NB, npad)
            A[lo, blk(k)] = A[lo, blk(k)] @ X[blk(k), blk(k)].T                 # panel: FP64 (K = 128, stays on DMMA)
            A[lo, lo] -= gemm(A[lo, blk(k)], A[lo, blk(k)])                     # trailing update
    Lm = np.tril(A)
    h = 1
    while h < nb:                                                                 # trtri by recursive doubling
        for g0 in range(0, nb, 2 * h):
            top = slice(g0 * NB, min(g0 + h, nb) * NB)
            bot = slice(min(g0 + h, nb) * NB, min(g0 + 2 * h, nb) * NB)
            if bot.start >= bot.stop:
                continue
            T = gemm(Lm[bot, top], X[top, top].T)                                # L21 X11
            X[bot, top] = -gemm(X[bot, bot], T.T)                                # - X22 (L21 X11)
        h *= 2
    J = gemm(X.T, X.T)                                                           # lauum: X' X
    return J[:n, :n]


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
    orc = g.load_oracle()
    orc.build()
    Z = orc.synth_alignment(L, M, 20140321)
    q = int(Z.max())
    Pi_t, Pij_t, Meff, W, info = orc.compute_weighted_frequencies(Z, q, "auto")
    print(f"L={L} M={M} n={(q - 1) * L} theta={info['theta']:.4f} Meff={Meff:.1f}")
    for pc in (0.8, 0.2):
        C = orc.compute_C(*orc.add_pseudocount(Pi_t, Pij_t, pc, q))
        ref = np.linalg.inv(C)
        cond = np.linalg.cond(C)
        base = blocked_inverse(C, lambda a, b: a @ b.T)
        nrm = np.max(np.abs(ref))
        print(f"pc={pc}: cond(C)={cond:.3g}; blocked FP64 vs LAPACK inv: {np.max(np.abs(base - ref)) / nrm:.2e}")
        for w, s in ((7, 4), (7, 5), (7, 6), (7, 7), (7, 8), (6, 7), (6, 8)):
            t0 = time.time()
            J = blocked_inverse(C, lambda a, b: sliced_gemm(a, b, w, s))
            err = np.max(np.abs(J - ref)) / nrm
            print(f"   w={w} s={s} ({s * (s + 1) // 2:2d} int8 GEMMs per FP64 GEMM): normwise error {err:.2e}   [{time.time() - t0:.1f} s]")


if __name__ == "__main__":
    main()
