"""The INT8-sliced FP64 GEMM (csrc/ozaki.cu) that carries the big products of mJ = inv(cholesky(C)) (reference
src/GaussDCA.jl:34) on the tcgen05 tensor cores.

Checked three ways:
  * bit for bit against an integer model of the same arithmetic in numpy (exact digit split, exact int64 digit products,
    one FP64 rounding) -- the kernel's S32 accumulation and its recombination are exact, so every output bit is determined;
  * against the plain FP64 product, every operand orientation and triangular k-range flag the inversion uses;
  * the whole inversion with the engine on and off: mJ within 1e-11 normwise of the DMMA path, 1e-9 of LAPACK.
"""
import ctypes

import numpy as np
import pytest

from test_gpu_parity import TOL, assert_rank_equal_tie_aware, normwise

pytestmark = pytest.mark.gpu

S, W = 8, 7


def gemm(ctx, engine, A, a_cols, B, b_cols, C, flags=0, alpha=1.0, beta=0.0):
    from gaussdca_jl_b200._lib import ptr
    m = A.shape[1] if a_cols else A.shape[0]
    k = A.shape[0] if a_cols else A.shape[1]
    n = B.shape[1] if b_cols else B.shape[0]
    assert (B.shape[0] if b_cols else B.shape[1]) == k and C.shape == (m, n)
    A, B = np.ascontiguousarray(A), np.ascontiguousarray(B)
    out = np.ascontiguousarray(C).copy()
    ctx.check(ctx.lib.gdca_test_fp64_gemm(ctx.h, engine, ptr(A), int(a_cols), ptr(B), int(b_cols), ptr(out), m, n, k, flags,
                                          float(alpha), float(beta)))
    return out


def slice_rows(A):
    """rows of A -> (scale 2^e, digits int64 [S][rows][k]) with A = 2^e sum_t d_t 2^(-7 (t+1)) + tiny, exactly as ozaki.cu."""
    mx = np.max(np.abs(A), axis=1)
    e = np.where(mx > 0, np.frexp(mx)[1] + 1, 0)
    rem = np.ldexp(A, -e[:, None])
    D = []
    for t in range(S):
        sc = 2.0 ** (W * (t + 1))
        d = np.rint(rem * sc)
        rem = rem - d / sc
        D.append(d.astype(np.int64))
    return np.ldexp(1.0, e), D


def digit_model(A, B, alpha=1.0):
    """alpha A B^T in the kernel's arithmetic: exact integers, two exact conversions, one rounding."""
    sa, DA = slice_rows(A)
    sb, DB = slice_rows(B)
    Dd = [sum(DA[t] @ DB[d - t].T for t in range(d + 1)) for d in range(S)]
    hi = (Dd[0] << 21) + (Dd[1] << 14) + (Dd[2] << 7) + Dd[3]
    lo = (Dd[4] << 21) + (Dd[5] << 14) + (Dd[6] << 7) + Dd[7]
    v = hi.astype(np.float64) * 2.0 ** -35 + lo.astype(np.float64) * 2.0 ** -63
    return (alpha * sa)[:, None] * v * sb[None, :]


def spread(rng, shape):
    """entries with a spread of magnitudes inside each row (what rows of L and inv(L) look like)"""
    return rng.standard_normal(shape) * np.exp2(rng.integers(-12, 4, size=shape).astype(np.float64))


@pytest.mark.parametrize("m,n,k", [(128, 128, 128), (256, 128, 384), (384, 256, 1024)])
def test_sliced_gemm_is_bit_exact_against_the_integer_model(ctx, m, n, k):
    rng = np.random.default_rng(m + n + k)
    A, B = spread(rng, (m, k)), spread(rng, (n, k))
    A[3] = 0.0                                     # an all-zero row: exponent 0, all digits 0
    B[5, :] = 2.0 ** -300 * rng.standard_normal(k)  # a tiny row scales like any other
    C0 = np.zeros((m, n))
    got = gemm(ctx, 1, A, 0, B, 0, C0)
    want = digit_model(A, B)
    assert np.array_equal(got, want)
    # accumulate form, negative alpha (the trailing update of the Cholesky)
    C1 = rng.standard_normal((m, n))
    got = gemm(ctx, 1, A, 0, B, 0, C1, alpha=-1.0, beta=1.0)
    assert np.array_equal(got, C1 + digit_model(A, B, -1.0))
    # and the split loses less than an FP64 dot product does: 2^-56 of the row scales per term
    ref = A @ B.T
    bound = np.max(np.abs(A), axis=1)[:, None] * np.max(np.abs(B), axis=1)[None, :] * k
    assert np.max(np.abs(want - ref) / np.maximum(bound, 1e-300)) <= 2.0 ** -52


@pytest.mark.parametrize("a_cols,b_cols", [(0, 0), (0, 1), (1, 1), (1, 0)])
def test_operand_orientations_match_fp64(ctx, a_cols, b_cols):
    rng = np.random.default_rng(10 * a_cols + b_cols)
    m, n, k = 384, 256, 640
    A = spread(rng, (k, m) if a_cols else (m, k))
    B = spread(rng, (k, n) if b_cols else (n, k))
    ref = (A.T if a_cols else A) @ (B if b_cols else B.T)
    got = gemm(ctx, 1, A, a_cols, B, b_cols, np.zeros((m, n)))
    Ar, Br = (A.T if a_cols else A), (B.T if b_cols else B)
    assert np.array_equal(got, digit_model(Ar, Br))          # the transposed slicer produces the same digits
    bound = np.max(np.abs(Ar), axis=1)[:, None] * np.max(np.abs(Br), axis=1)[None, :] * k
    assert np.max(np.abs(got - ref) / bound) <= 2.0 ** -52
    if not (a_cols and not b_cols):                            # the DMMA kernel has no such instantiation
        assert normwise(gemm(ctx, 0, A, a_cols, B, b_cols, np.zeros((m, n))), ref) <= 1e-14


def test_triangular_k_ranges_of_the_inversion(ctx):
    """The three restricted k ranges chol.cu uses (trtri level products, lauum), both engines, against the full product of
    operands that really are triangular."""
    rng = np.random.default_rng(7)
    m, n, k = 512, 384, 512
    tol = {0: 1e-14, 1: 1e-13}   # sliced engine: 2^-56 of (row scale x column scale) per term, rows here span 16 binades
    # flag 2: B lower triangular as [k][n]  (T = L21 * X11)
    A = spread(rng, (m, k))
    Bkn = np.tril(spread(rng, (k, k)))[:, :n]
    ref = A @ Bkn
    for eng in (0, 1):
        assert normwise(gemm(ctx, eng, A, 0, Bkn, 1, np.zeros((m, n)), flags=2), ref) <= tol[eng]
    # flag 8: A lower triangular as [m][k]  (X21 = -X22 * T), alpha = -1
    Amk = np.tril(spread(rng, (m, m)))
    Bkn2 = spread(rng, (m, n))
    ref = -Amk @ Bkn2
    for eng in (0, 1):
        assert normwise(gemm(ctx, eng, Amk, 0, Bkn2, 1, np.zeros((m, n)), flags=8, alpha=-1.0), ref) <= tol[eng]
    # flags 1 | 4: lauum  J = X' X, X lower triangular, output tiles on and below the diagonal only
    X = np.tril(spread(rng, (k, k)))
    ref = X.T @ X
    for eng in (0, 1):
        got = gemm(ctx, eng, X, 1, X, 1, np.full((k, k), 7.0), flags=1 | 4)
        low = np.tril(np.ones((k, k), dtype=bool))
        assert np.max(np.abs(got - ref)[low]) / np.max(np.abs(ref)) <= tol[eng]
        # tiles strictly above the block diagonal are not touched (128-row tiles of the DMMA kernel, 128 x 64 of the sliced one)
        assert np.all(got[:128, 128:] == 7.0)


@pytest.mark.parametrize("n", [2048, 2500, 4100])
def test_inverse_sliced_engine_vs_dmma_and_lapack(pkg, orc, ctx, n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n + 10))
    C = A @ A.T / (n + 10) + 0.05 * np.eye(n)
    lib = ctx.lib
    try:
        ctx.check(lib.gdca_set_ozaki(ctx.h, 0))
        mJ0 = pkg.inverse(C, ctx=ctx)
        on = ctypes.c_int32()
        ctx.check(lib.gdca_dev_inverse_info(ctx.h, ctypes.byref(on), None, None))
        assert on.value == 0
        ctx.check(lib.gdca_set_ozaki(ctx.h, 1))
        mJ1 = pkg.inverse(C, ctx=ctx)
        ops, flop = ctypes.c_double(), ctypes.c_double()
        ctx.check(lib.gdca_dev_inverse_info(ctx.h, ctypes.byref(on), ctypes.byref(ops), ctypes.byref(flop)))
        assert on.value == 1 and ops.value > 0 and flop.value > 0.5 * n ** 3   # most of the n^3 flop went through INT8
    finally:
        ctx.check(lib.gdca_set_ozaki(ctx.h, 1))
    ref = orc.inv_cholesky(C)
    assert np.array_equal(mJ1, mJ1.T)
    assert normwise(mJ1, mJ0) <= 1e-11
    assert normwise(mJ1, ref) <= TOL and normwise(mJ0, ref) <= TOL
    assert np.max(np.abs(mJ1 @ C - np.eye(n))) <= 1e-9


def test_inverse_sliced_engine_on_an_oracle_covariance_and_not_spd(pkg, orc, ctx):
    """A real covariance (n = 4000, pseudocount 0.2: cond ~ 1e3..1e4), both engines; and LAPACK's info on a failed pivot that
    sits behind sliced trailing updates."""
    Z = orc.synth_alignment(200, 3000, seed=5)
    q = int(Z.max())
    counts, Wt, Meff, _ = orc.compute_weights(Z, 0.3)
    Pi_t, Pij_t = orc.compute_freqs(Z, q, Wt, Meff)
    C = orc.compute_C(*orc.add_pseudocount(Pi_t, Pij_t, 0.2, q))
    ref = orc.inv_cholesky(C)
    mJ = pkg.inverse(C, ctx=ctx)
    assert normwise(mJ, ref) <= TOL
    S1 = pkg.compute_FN(mJ, q, ctx=ctx)
    assert normwise(S1, orc.compute_FN(ref, q)) <= TOL
    Cb = C.copy()
    Cb[3000:, :] = 0
    Cb[:, 3000:] = 0
    with pytest.raises(pkg.PosDefException) as ei:
        pkg.inverse(Cb, ctx=ctx)
    assert ei.value.info == 3001


def test_end_to_end_ranking_same_with_both_engines(pkg, orc, ctx):
    Z = orc.synth_alignment(200, 6000, seed=20140321)
    try:
        ctx.check(ctx.lib.gdca_set_ozaki(ctx.h, 0))
        R0 = pkg.gdca_from_alignment(Z, ctx=ctx)
    finally:
        ctx.check(ctx.lib.gdca_set_ozaki(ctx.h, 1))
    R1 = pkg.gdca_from_alignment(Z, ctx=ctx)
    assert_rank_equal_tie_aware(R1, R0)
    assert [(i, j) for i, j, _ in R1[:200]] == [(i, j) for i, j, _ in R0[:200]]
