"""Tile order of the tensor-core prefilter (csrc/tcfilter.cu), replayed on the host through the C ABI -- the kernel's own
iterator compiled for the host, no GPU needed: every 128 x 128 block (bi <= bj) of the pair matrix is covered, each tile is
visited by exactly one (rank, CTA), and the two CTAs of a cluster walk neighbouring rows of the same column tile in lock step."""
import ctypes
import itertools

import numpy as np
import pytest


def order(lib, T, bits, rank, world, grid, cta):
    from gaussdca_jl_b200._lib import ptr
    n = ctypes.c_int64()
    assert lib.gdca_tc_filter_tile_order(T, bits, rank, world, grid, cta, None, 0, ctypes.byref(n)) == 0
    out = np.zeros((max(1, n.value), 4), dtype=np.int32)
    assert lib.gdca_tc_filter_tile_order(T, bits, rank, world, grid, cta, ptr(out), n.value, ctypes.byref(n)) == 0
    return out[: n.value]


@pytest.fixture(scope="module")
def lib(pkg):
    from gaussdca_jl_b200 import _lib
    return _lib.load()


@pytest.mark.parametrize("T,bits,world,grid", [(1, 4, 1, 148), (2, 8, 1, 148), (7, 4, 1, 4), (37, 4, 1, 148), (37, 8, 3, 148),
                                              (100, 4, 2, 148), (163, 8, 1, 148), (163, 4, 8, 148), (1563, 4, 1, 148),
                                              (1563, 8, 2, 148), (400, 4, 5, 16)])
def test_tiles_cover_the_upper_triangle_exactly_once(lib, T, bits, world, grid):
    colw = 224 if bits == 4 else 256
    NT = -(-T * 128 // colw)
    seen = {}
    for rank in range(world):
        for cta in range(grid):
            rows = order(lib, T, bits, rank, world, grid, cta)
            for bi, cj, valid, _ in rows.tolist():
                if valid:
                    assert bi % world == rank and 0 <= bi < T and 0 <= cj < NT
                    assert (bi, cj) not in seen, (bi, cj, seen[(bi, cj)], (rank, cta))
                    seen[(bi, cj)] = (rank, cta)
    # exactly the tiles that reach the diagonal block of their row or lie to the right of it
    want = {(bi, cj) for bi in range(T) for cj in range(NT) if colw * (cj + 1) > 128 * bi}
    assert set(seen) == want
    # hence every block (bi <= bj) is covered by the tiles of ONE rank (whole rows per rank)
    for bi in range(T):
        for bj in range(bi, T):
            cjs = range(128 * bj // colw, (128 * bj + 127) // colw + 1)
            assert all((bi, cj) in seen for cj in cjs)


@pytest.mark.parametrize("T,bits,world", [(37, 4, 1), (163, 4, 2), (163, 8, 3), (1563, 4, 1)])
def test_cluster_pairs_walk_in_lock_step(lib, T, bits, world):
    grid = 148
    for rank, c in itertools.product(range(world), (0, 1, 36, 73)):
        a = order(lib, T, bits, rank, world, grid, 2 * c)
        b = order(lib, T, bits, rank, world, grid, 2 * c + 1)
        assert len(a) == len(b)
        assert np.array_equal(a[:, 1], b[:, 1])                    # same column tile
        assert np.array_equal(a[:, 2], b[:, 3]) and np.array_equal(a[:, 3], b[:, 2])   # each sees the other's validity
        assert np.array_equal(b[:, 0] - a[:, 0], np.full(len(a), world))               # neighbouring rows of the rank


def test_projection_bound_is_conservative_for_every_alphabet():
    """The exactness argument of the prefilter, checked in integers on the CPU: with classes = state & 3 and the simplex code,
    S = 4 ident_proj - L and hamming_proj = (3L - S)/4 <= hamming for EVERY pair, for every alphabet size q <= 31; hence a
    neighbour pair (hamming < thresh) always has S > 3L - 4 thresh and its cell can never be cleared."""
    rng = np.random.default_rng(11)
    code = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]])     # v_a . v_b = 3 if a == b else -1
    assert np.array_equal(code @ code.T, 4 * np.eye(4, dtype=int) - 1)
    for q in (2, 4, 5, 21, 31):
        for L in (1, 7, 53, 200):
            M = 60
            base = rng.integers(1, q + 1, size=(6, L))
            Z = base[rng.integers(0, 6, M)].copy()
            mut = rng.random((M, L)) < rng.random((M, 1))                     # per-sequence mutation rates 0..1
            Z[mut] = rng.integers(1, q + 1, size=int(mut.sum()))
            V = code[Z & 3].reshape(M, 3 * L)                                 # what encode_simplex*_kernel writes
            S = V @ V.T
            ham = (Z[:, None, :] != Z[None, :, :]).sum(-1)
            ident_proj = ((Z[:, None, :] & 3) == (Z[None, :, :] & 3)).sum(-1)
            assert np.array_equal(S, 4 * ident_proj - L)
            ham_proj = (3 * L - S) // 4
            assert np.array_equal(4 * ham_proj, 3 * L - S) and np.all(ham_proj <= ham)
            for thresh in (0, 1, L // 4, L // 2, L):
                assert np.all(S[ham < thresh] > 3 * L - 4 * thresh)
