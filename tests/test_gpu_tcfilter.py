"""Tensor-core prefilter of the neighbour-count sweep (csrc/tcfilter.cu): tcgen05 scores are exact integers, the
flagged blocks are a superset of the blocks holding a neighbour pair, and counts / weights are bit-identical with
the filter off, on, and against the CPU oracle (DCAUtils compute_weights, call site src/GaussDCA.jl:28)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def projected_scores(Z):
    """S[k,l] = 4 * #{i : class(Z[k,i]) == class(Z[l,i])} - L, class = state & 3 (exact integers)."""
    cls = (Z.astype(np.int64) & 3)
    M, L = Z.shape
    ident = np.zeros((M, M), dtype=np.int64)
    for c in range(4):
        X = (cls == c).astype(np.float64)
        ident += np.rint(X @ X.T).astype(np.int64)
    return 4 * ident - L


# operands: e4m3 (kind::f8f6f4) / packed e2m1 (kind::mxf4, unit block scales) / int8 (kind::i8, code 80); launch: independent CTAs / 2-CTA clusters
# sharing the column tile by TMA multicast / 2-CTA pairs issuing one cta_group::2 MMA of 256 x 224 (each CTA keeps half of the column tile)
VARIANTS = [(8, 0), (4, 0), (8, 1), (4, 1), (80, 1), (4, 2)]


@pytest.fixture(params=VARIANTS, ids=["fp8", "fp4", "fp8-multicast", "fp4-multicast", "int8-multicast", "fp4-cta_group2"])
def bits(request, ctx):
    b, mc = request.param
    ctx.check(ctx.lib.gdca_set_tc_filter_bits(ctx.h, b))
    ctx.check(ctx.lib.gdca_set_tc_filter_multicast(ctx.h, mc))
    yield b
    ctx.check(ctx.lib.gdca_set_tc_filter_bits(ctx.h, 4))
    ctx.check(ctx.lib.gdca_set_tc_filter_multicast(ctx.h, 2))


def run_filter(ctx, Z, thresh, want_scores=True):
    from gaussdca_jl_b200._lib import ptr
    lib = ctx.lib
    M, L = Z.shape
    ctx.check(lib.gdca_dev_load(ctx.h, ptr(Z), L, M))
    T = (M + 127) // 128
    ld = T * 128 + 256
    flags = np.zeros(T * T, dtype=np.uint32)
    S = np.zeros((T * 128, ld), dtype=np.float32) if want_scores else None
    ctx.check(lib.gdca_dev_tc_filter(ctx.h, thresh, ptr(flags), ptr(S) if want_scores else None, ld))
    return flags.reshape(T, T), S, T


@pytest.mark.parametrize("L,M", [(53, 300), (128, 512), (200, 1000), (43, 129), (500, 700), (342, 1500)])
def test_filter_scores_and_flags_are_exact(orc, ctx, bits, L, M):
    Z = orc.synth_alignment(L, M, seed=11 + L + M)
    thresh = int(0.4 * L)
    flags, S, T = run_filter(ctx, Z, thresh)
    want = projected_scores(Z)
    bound = 3 * L - 4 * thresh
    for bi in range(T):
        for bj in range(bi, T):
            r0, r1, c0, c1 = bi * 128, min(M, bi * 128 + 128), bj * 128, min(M, bj * 128 + 128)
            got = S[r0:r1, c0:c1]
            assert np.array_equal(got, want[r0:r1, c0:c1].astype(np.float32)), (bits, bi, bj)
            # padding rows / columns are all-zero vectors: S = 0 there
            blk = np.zeros((128, 128))
            blk[: r1 - r0, : c1 - c0] = want[r0:r1, c0:c1]
            mask = 0
            for cr in range(4):
                for cc in range(4):
                    if blk[32 * cr:32 * cr + 32, 32 * cc:32 * cc + 32].max() > bound:
                        mask |= 1 << (4 * cr + cc)
            assert flags[bi, bj] == mask, (bits, bi, bj, hex(int(flags[bi, bj])), hex(mask))
    assert not np.any(np.tril(flags, -1))


def test_flagged_blocks_cover_every_neighbour_pair(orc, ctx, bits):
    rng = np.random.default_rng(5)
    L, M = 160, 2000
    Z = orc.synth_alignment(L, M, seed=99)
    for thresh in (1, 20, 64, 80, 159):
        flags, _, T = run_filter(ctx, Z, thresh, want_scores=False)
        ham = np.zeros((M, M), dtype=np.int64)
        for c in range(1, 32):
            X = (Z == c).astype(np.float64)
            ham += np.rint(X @ X.T).astype(np.int64)
        ham = L - ham
        nb = ham < thresh
        rr, cc = np.nonzero(np.triu(nb, 1))              # every neighbour pair lies in a flagged cell
        bit = 4 * ((rr % 128) // 32) + (cc % 128) // 32
        ok = (flags[rr // 128, cc // 128].astype(np.int64) >> bit) & 1
        assert ok.all(), (thresh, int((ok == 0).sum()))
    del rng


@pytest.mark.parametrize("L,M", [(53, 300), (200, 1000), (97, 2500), (150, 6000)])
def test_counts_identical_with_and_without_filter(pkg, orc, ctx, bits, L, M):
    Z = orc.synth_alignment(L, M, seed=1 + L + M)
    lib = ctx.lib
    try:
        for theta in ("auto", 0.3, 0.12, 1.0):
            tho = orc.compute_theta(Z) if theta == "auto" else theta
            counts, W, Meff, thresh = orc.compute_weights(Z, tho)
            ctx.check(lib.gdca_set_tc_filter(ctx.h, 0))
            w0 = pkg.compute_weights(Z, theta, ctx=ctx, full=True)
            ctx.check(lib.gdca_set_tc_filter(ctx.h, 2))
            w2 = pkg.compute_weights(Z, theta, ctx=ctx, full=True)
            filt = ctypes.c_int32()
            ctx.check(lib.gdca_dev_sweep_info(ctx.h, ctypes.byref(filt), None, None, None, None, None, None))
            assert filt.value == bits
            for w in (w0, w2):
                assert w["thresh"] == thresh
                assert np.array_equal(w["counts"], counts), (L, M, theta)
                assert np.array_equal(w["W"], W)
                assert w["Meff"] == Meff
    finally:
        ctx.check(lib.gdca_set_tc_filter(ctx.h, 1))


def test_filtered_sweep_adversarial_and_sharded(pkg, orc, ctx, bits):
    """Neighbour pairs spread over many blocks, near-threshold pairs, and the multi-GPU partition run sequentially:
    the filter's per-rank tile share plus the exact sweep of its flagged blocks adds up to the unsharded counts."""
    from gaussdca_jl_b200._lib import ptr
    rng = np.random.default_rng(17)
    L, M = 96, 3000
    base = rng.integers(1, 22, size=(40, L), dtype=np.int8)
    Z = base[rng.integers(0, 40, size=M)].copy()
    nmut = rng.integers(0, 60, size=M)
    for k in range(M):
        pos = rng.choice(L, size=nmut[k], replace=False)
        Z[k, pos] = rng.integers(1, 22, size=nmut[k])
    Z[0, 0] = 21
    lib = ctx.lib
    thresh = 30
    want = orc.compute_weights(Z, thresh / L + 1e-9)[0]

    def grab():
        c = np.empty(M, dtype=np.int32)
        ctx.check(lib.gdca_dev_copy_to_host(ctx.h, ptr(c), lib.gdca_dev_counts_ptr(ctx.h), M * 4))
        return c

    try:
        ctx.check(lib.gdca_set_tc_filter(ctx.h, 2))
        ctx.check(lib.gdca_dev_load(ctx.h, ptr(Z), L, M))
        ctx.check(lib.gdca_dev_pair_pass(ctx.h, 1, thresh))
        assert np.array_equal(grab() + 1, want)
        for world in (2, 3):
            tot = np.zeros(M, dtype=np.int32)
            for r in range(world):
                ctx.check(lib.gdca_set_shard(ctx.h, r, world))
                ctx.check(lib.gdca_dev_pair_pass(ctx.h, 1, thresh))
                tot += grab()
            ctx.check(lib.gdca_set_shard(ctx.h, 0, 1))
            assert np.array_equal(tot + 1, want), world
    finally:
        ctx.check(lib.gdca_set_shard(ctx.h, 0, 1))
        ctx.check(lib.gdca_set_tc_filter(ctx.h, 1))


def test_large_run_same_ranking_with_and_without_filter(pkg, orc, ctx, bits):
    """Config-B-sized sweep (auto mode switches the filter on at M >= 16384): identical weights and ranking."""
    L, M = 64, 20000
    Z = orc.synth_alignment(L, M, seed=20140321)
    lib = ctx.lib
    try:
        ctx.check(lib.gdca_set_tc_filter(ctx.h, 0))
        w0 = pkg.compute_weights(Z, "auto", ctx=ctx, full=True)
        R0 = pkg.gdca_from_alignment(Z, 0.8, "auto", "frob", 5, ctx=ctx)
        ctx.check(lib.gdca_set_tc_filter(ctx.h, 1))
        w1 = pkg.compute_weights(Z, "auto", ctx=ctx, full=True)
        filt, blocks = ctypes.c_int32(), ctypes.c_int64()
        ctx.check(lib.gdca_dev_sweep_info(ctx.h, ctypes.byref(filt), None, None, ctypes.byref(blocks), None, None, None))
        assert filt.value == bits
        T = (M + 127) // 128
        assert 0 < blocks.value <= T * (T + 1) // 2
        R1 = pkg.gdca_from_alignment(Z, 0.8, "auto", "frob", 5, ctx=ctx)
    finally:
        ctx.check(lib.gdca_set_tc_filter(ctx.h, 1))
    assert np.array_equal(w0["counts"], w1["counts"])
    assert w0["Meff"] == w1["Meff"]
    assert R0 == R1
