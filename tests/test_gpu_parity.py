"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference's goldens.

Tolerances (BASELINE.json north_star): neighbour counts / thresh / W bit-exact, Meff equal to the
correctly rounded value; C, mJ, raw scores, APC scores: normwise max|d| / max|.| <= 1e-9 in FP64;
ranking identical up to mathematical ties (scores within 1e-9 * max|S|)."""
import io
import math

import numpy as np
import pytest

from conftest import GOLDEN_CASES, golden_path, printed_todict, rank_todict, read_golden

pytestmark = pytest.mark.gpu

TOL = 1e-9


def normwise(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def load_case(pkg, fa, kw):
    Z = pkg.read_fasta_alignment(golden_path(fa), kw.get("max_gap_fraction", 0.9))
    if kw.get("remove_dups"):
        Z, _ = pkg.remove_duplicate_sequences(Z)
    return Z


# ----------------------------------------------------------------------------- goldens, end to end
@pytest.mark.parametrize("name,fa,kw", GOLDEN_CASES, ids=[c[0] for c in GOLDEN_CASES])
def test_golden_end_to_end(pkg, ctx, name, fa, kw):
    """test/runtests.jl:52-76 with the same comparison: key sets equal, values equal after %e printing."""
    R = pkg.gDCA(golden_path(fa), ctx=ctx, **kw)
    buf = io.StringIO()
    pkg.printrank(buf, R)
    got = printed_todict(buf.getvalue())
    want = read_golden(name)
    assert sorted(got) == sorted(want)
    # Julia's isapprox default rtol = sqrt(eps) on the 7-digit printed values
    bad = [k for k in want if not math.isclose(got[k], want[k], rel_tol=1.5e-8 + 1.01e-6, abs_tol=0.0)]
    assert not bad, (len(bad), bad[:5], [(got[k], want[k]) for k in bad[:5]])
    # ranking is sorted descending
    xs = [x for _, _, x in R]
    assert all(xs[t] >= xs[t + 1] for t in range(len(xs) - 1))


@pytest.mark.parametrize("name,fa,kw", GOLDEN_CASES, ids=[c[0] for c in GOLDEN_CASES])
def test_stages_vs_oracle_on_fixtures(pkg, orc, ctx, name, fa, kw):
    Z = load_case(pkg, fa, kw)
    theta = kw.get("theta", "auto")
    pc = kw.get("pseudocount", 0.8)
    score = kw.get("score", "frob")
    ms = kw.get("min_separation", 5)
    st = {}
    Ro = orc.gdca_from_Z(Z, pc, theta, score, ms, stages=st)
    w = pkg.compute_weights(Z, theta, ctx=ctx, full=True)
    assert np.array_equal(w["counts"], st["counts"])
    assert np.array_equal(w["W"], st["W"])            # bit-exact doubles
    assert w["thresh"] == st["thresh"]
    assert w["theta"] == st["theta"]
    assert w["Meff"] == st["Meff"]
    C, Pi, q = pkg.compute_covariance(Z, w["W"], w["Meff"], pc, ctx=ctx)
    assert q == st["q"]
    assert np.array_equal(C, C.T)
    assert normwise(C, st["C"]) <= 1e-13
    mJ = pkg.inverse(C, ctx=ctx)
    assert np.array_equal(mJ, mJ.T)
    assert normwise(mJ, st["mJ"]) <= TOL
    S = pkg.compute_DI_gauss(mJ, C, q, ctx=ctx) if score == "DI" else pkg.compute_FN(mJ, q, ctx=ctx)
    assert np.array_equal(S, S.T) and np.all(np.diag(S) == 0)
    assert normwise(S, st["S_raw"]) <= TOL
    # stage isolation: feed the oracle's mJ, compare the score kernel alone
    S_iso = (pkg.compute_DI_gauss(st["mJ"], st["C"], q, ctx=ctx) if score == "DI" else pkg.compute_FN(st["mJ"], q, ctx=ctx))
    assert normwise(S_iso, st["S_raw"]) <= 1e-12
    Sc = pkg.correct_APC(st["S_raw"], ctx=ctx)
    assert normwise(Sc, st["S"]) <= 1e-13
    R = pkg.compute_ranking(st["S"], ms, ctx=ctx)
    assert R == Ro                                     # same input -> identical order incl. ties, same bits
    # fused run: tie-aware ranking equality
    Rf = pkg.gdca_from_alignment(Z, pc, theta, score, ms, ctx=ctx)
    assert_rank_equal_tie_aware(Rf, Ro)


def assert_rank_equal_tie_aware(R, Ro, tol=TOL):
    d, do = rank_todict(R), rank_todict(Ro)
    assert sorted(d) == sorted(do)
    smax = max(abs(x) for x in do.values())
    assert max(abs(d[k] - do[k]) for k in do) <= tol * smax
    # positions may differ only inside runs of (near-)tied scores
    pos_o = {(i, j): t for t, (i, j, _) in enumerate(Ro)}
    for t, (i, j, x) in enumerate(R):
        to = pos_o[(i, j)]
        if to != t:
            assert abs(Ro[t][2] - do[(i, j)]) <= 4 * tol * smax, (t, to, (i, j), Ro[t], x)


# ----------------------------------------------------------------------------- weights: exact integers
@pytest.mark.parametrize("L,M", [(1, 2), (31, 5), (32, 127), (33, 128), (64, 129), (53, 300), (200, 1000), (97, 2500)])
def test_weights_bitexact_synthetic(pkg, orc, ctx, L, M):
    Z = orc.synth_alignment(L, M, seed=7 + L + M)
    for theta in ("auto", 0.3, 0.0, 1.0):
        if theta == "auto":
            tho = orc.compute_theta(Z)
        else:
            tho = theta
        counts, W, Meff, thresh = orc.compute_weights(Z, tho)
        w = pkg.compute_weights(Z, theta, ctx=ctx, full=True)
        assert w["theta"] == tho
        assert w["thresh"] == thresh
        assert np.array_equal(w["counts"], counts), (L, M, theta)
        assert np.array_equal(w["W"], W)
        assert w["Meff"] == Meff


def test_weights_bytes_and_packed_oracles_agree_with_gpu(pkg, orc, ctx):
    """The reference's fast/fallback duality (test/runtests.jl:78-86): both CPU paths and the GPU agree."""
    Z = orc.synth_alignment(77, 700, seed=3)
    th = orc.compute_theta(Z, packed=True)
    assert th == orc.compute_theta(Z, packed=False)
    c1 = orc.compute_weights(Z, th, packed=True)[0]
    c2 = orc.compute_weights(Z, th, packed=False)[0]
    w = pkg.compute_weights(Z, "auto", ctx=ctx, full=True)
    assert np.array_equal(c1, c2) and np.array_equal(c1, w["counts"])


def test_weights_duplicates_and_small_q(pkg, orc, ctx):
    rng = np.random.default_rng(5)
    Z = rng.integers(1, 4, size=(200, 40), dtype=np.int8)  # q = 3 -> 2 bit planes
    Z[50:100] = Z[0:50]                                     # exact duplicates: hamming 0
    counts, W, Meff, thresh = orc.compute_weights(Z, 0.2)
    w = pkg.compute_weights(Z, 0.2, ctx=ctx, full=True)
    assert np.array_equal(w["counts"], counts) and w["Meff"] == Meff
    assert counts[:100].min() >= 2


def test_theta_auto_needs_one_sweep_and_is_exact(pkg, orc, ctx):
    """theta=:auto comes from per-site state histograms (no sweep); the count sweep exits early -- still bit-exact."""
    Z = orc.synth_alignment(100, 20000, seed=11)
    tho = orc.compute_theta(Z)
    counts, W, Meff, thresh = orc.compute_weights(Z, tho)
    w = pkg.compute_weights(Z, "auto", ctx=ctx, full=True)
    assert w["theta"] == tho and w["thresh"] == thresh
    assert np.array_equal(w["counts"], counts) and w["Meff"] == Meff
    assert w["passes"] == 1


def test_sharded_pair_sweep_sums_to_unsharded(pkg, orc, ctx):
    """Multi-GPU partition run sequentially on one GPU: integer partials add up bit-exactly."""
    Z = orc.synth_alignment(90, 3000, seed=21)
    M, L = Z.shape
    lib = ctx.lib
    from gaussdca_jl_b200._lib import ptr
    ctx.check(lib.gdca_dev_load(ctx.h, ptr(Z), L, M))
    stride = lib.gdca_dev_counts_stride(ctx.h)

    def grab():
        c = np.empty(M, dtype=np.int32)
        h = np.empty(2, dtype=np.int64)
        ctx.check(lib.gdca_dev_copy_to_host(ctx.h, ptr(c), lib.gdca_dev_counts_ptr(ctx.h), M * 4))
        ctx.check(lib.gdca_dev_copy_to_host(ctx.h, ptr(h), lib.gdca_dev_ham_sum_ptr(ctx.h), 16))
        return c, h

    thresh = 30
    ctx.check(lib.gdca_set_shard(ctx.h, 0, 1))
    ctx.check(lib.gdca_dev_pair_pass(ctx.h, 2, thresh))
    c_full, h_full = grab()
    world = 3
    c_sum, h_sum = np.zeros_like(c_full), np.zeros_like(h_full)
    for r in range(world):
        ctx.check(lib.gdca_set_shard(ctx.h, r, world))
        ctx.check(lib.gdca_dev_pair_pass(ctx.h, 2, thresh))
        c, h = grab()
        c_sum += c
        h_sum += h
    ctx.check(lib.gdca_set_shard(ctx.h, 0, 1))
    assert np.array_equal(c_sum, c_full) and np.array_equal(h_sum, h_full)
    assert h_full[1] == M * (M - 1) // 2
    counts = orc.compute_weights(Z, thresh / L + 1e-9)[0]  # floor(theta*L) == thresh
    # counts buffer row 1 of mode 2 holds `thresh` itself
    ctx.check(lib.gdca_dev_pair_pass(ctx.h, 1, thresh))
    c1, _ = grab()
    assert np.array_equal(c1 + 1, counts)
    assert orc.ident_sum(Z) == M * (M - 1) // 2 * L - int(h_full[0])
    # the O(M L) histogram route gives the same exact integer as the O(M^2 L) sweep
    import ctypes
    v = ctypes.c_uint64()
    ctx.check(lib.gdca_dev_ident_sum(ctx.h, ctypes.byref(v)))
    assert v.value == orc.ident_sum(Z)


def test_early_exit_is_exact_on_adversarial_layouts(pkg, orc, ctx):
    """Neighbours that only differ in the LAST sites, far pairs that only differ in the FIRST sites, thresholds at
    word boundaries: the per-warp early exit must never change a count."""
    rng = np.random.default_rng(3)
    L, M = 96, 700
    base = rng.integers(1, 22, size=(1, L), dtype=np.int8)
    Z = np.repeat(base, M, axis=0)
    for k in range(M):
        if k % 3 == 0:      # differ only at the end
            nd = rng.integers(0, 40)
            Z[k, L - nd:] = rng.integers(1, 22, size=nd)
        elif k % 3 == 1:    # differ only at the beginning
            nd = rng.integers(0, 70)
            Z[k, :nd] = rng.integers(1, 22, size=nd)
        else:
            Z[k] = rng.integers(1, 22, size=L)
    Z[0, 0] = 21
    for theta in (1 / 96 + 1e-9, 0.25, 32 / 96 + 1e-9, 33 / 96 + 1e-9, 0.5, 64 / 96 + 1e-9, 1.0):
        counts, W, Meff, thresh = orc.compute_weights(Z, theta)
        w = pkg.compute_weights(Z, theta, ctx=ctx, full=True)
        assert w["thresh"] == thresh and np.array_equal(w["counts"], counts), theta


# ----------------------------------------------------------------------------- covariance / inverse / scores
@pytest.mark.parametrize("L,M,pc", [(7, 50, 0.8), (40, 2000, 0.5), (64, 1000, 0.2), (130, 300, 0.8), (33, 5000, 1.0)])
def test_covariance_vs_oracle(pkg, orc, ctx, L, M, pc):
    Z = orc.synth_alignment(L, M, seed=L * M)
    q = int(Z.max())
    counts, W, Meff, _ = orc.compute_weights(Z, 0.3)
    Pi_t, Pij_t = orc.compute_freqs(Z, q, W, Meff)
    Pi_o, Pij_o = orc.add_pseudocount(Pi_t, Pij_t, pc, q)
    C_o = orc.compute_C(Pi_o, Pij_o)
    C, Pi, qq = pkg.compute_covariance(Z, W, Meff, pc, ctx=ctx)
    assert qq == q
    assert np.array_equal(C, C.T)
    assert normwise(Pi, Pi_o) <= 1e-14
    assert normwise(C, C_o) <= 1e-13


def test_covariance_q20_no_gap_state(pkg, orc, ctx):
    """q = max(Z) is data dependent: no 21 anywhere -> q = 20, state 20 is dropped."""
    Z = orc.synth_alignment(30, 400, seed=9)
    Z[Z == 21] = 20
    q = int(Z.max())
    assert q == 20
    W = np.ones(Z.shape[0])
    Pi_t, Pij_t = orc.compute_freqs(Z, q, W, float(Z.shape[0]))
    Pi_o, Pij_o = orc.add_pseudocount(Pi_t, Pij_t, 0.8, q)
    C, Pi, qq = pkg.compute_covariance(Z, W, float(Z.shape[0]), 0.8, ctx=ctx)
    assert qq == 20 and C.shape == (19 * 30, 19 * 30)
    assert normwise(C, orc.compute_C(Pi_o, Pij_o)) <= 1e-13


@pytest.mark.parametrize("n", [20, 128, 129, 700, 1060, 2500])
def test_inverse_vs_lapack(pkg, orc, ctx, n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n + 10))
    C = A @ A.T / (n + 10) + 0.05 * np.eye(n)
    mJ = pkg.inverse(C, ctx=ctx)
    ref = orc.inv_cholesky(C)
    assert np.array_equal(mJ, mJ.T)
    assert normwise(mJ, ref) <= TOL
    assert np.max(np.abs(mJ @ C - np.eye(n))) <= 1e-9


def test_inverse_not_spd_raises_posdef(pkg, orc, ctx):
    """pseudocount = 0 on a small alignment gives a singular C: the reference throws PosDefException."""
    n = 300
    rng = np.random.default_rng(0)
    A = rng.standard_normal((n, n))
    C = A @ A.T
    C[150:, :] = 0
    C[:, 150:] = 0
    with pytest.raises(pkg.PosDefException) as ei:
        pkg.inverse(C, ctx=ctx)
    with pytest.raises(orc.PosDefException) as eo:
        orc.inv_cholesky(C)
    assert ei.value.info == eo.value.info == 151


@pytest.mark.parametrize("q", [21, 20, 5, 25, 31, 3])
def test_scores_vs_oracle_random_spd(pkg, orc, ctx, q):
    s, L = q - 1, 12
    n = s * L
    rng = np.random.default_rng(q)
    A = rng.standard_normal((n, 3 * n))
    C = A @ A.T / (3 * n) + 0.1 * np.eye(n)
    mJ = orc.inv_cholesky(C)
    assert normwise(pkg.compute_FN(mJ, q, ctx=ctx), orc.compute_FN(mJ, q)) <= 1e-12
    ref = orc.compute_DI_gauss(mJ, C, q)
    di_ql = pkg.compute_DI_gauss(mJ, C, q, ctx=ctx)     # default engine: tridiagonalisation + implicit QL, one lane per site pair
    assert normwise(di_ql, ref) <= 1e-11
    ctx.set_di_engine(0)                                  # the one-sided Jacobi engine, kept as the cross-check
    try:
        di_jac = pkg.compute_DI_gauss(mJ, C, q, ctx=ctx)
    finally:
        ctx.set_di_engine(1)
    assert normwise(di_jac, ref) <= 1e-11
    assert normwise(di_ql, di_jac) <= 1e-12
    assert np.array_equal(di_ql, di_ql.T) and not np.diag(di_ql).any()


def test_di_engines_agree_on_ragged_site_counts(pkg, orc, ctx):
    """L not a multiple of the 32 pairs a warp takes (and of the 96 of a CTA), rank-deficient couplings, a zero coupling block."""
    q, s = 21, 20
    for L in (33, 97, 130):
        n = s * L
        rng = np.random.default_rng(L)
        A = rng.standard_normal((n, 2 * n))
        C = A @ A.T / (2 * n) + 0.05 * np.eye(n)
        mJ = orc.inv_cholesky(C)
        mJ[0:s, s:2 * s] = 0.0; mJ[s:2 * s, 0:s] = 0.0                       # zero block: all eigenvalues 0
        mJ[2 * s:3 * s, 5 * s:6 * s][:, s // 2:] = 0.0                       # rank-deficient block
        mJ[5 * s:6 * s, 2 * s:3 * s] = mJ[2 * s:3 * s, 5 * s:6 * s].T
        di_ql = pkg.compute_DI_gauss(mJ, C, q, ctx=ctx)
        ctx.set_di_engine(0)
        try:
            di_jac = pkg.compute_DI_gauss(mJ, C, q, ctx=ctx)
        finally:
            ctx.set_di_engine(1)
        assert normwise(di_ql, di_jac) <= 1e-12, L
        assert abs(di_ql[0, 1]) <= 1e-13 and abs(di_ql[1, 0]) <= 1e-13     # s/2 log(1/2) + 1/2 s log 2 = 0
        if L == 33:
            assert normwise(di_ql, orc.compute_DI_gauss(mJ, C, q)) <= 1e-11


def test_apc_and_ranking_ties_are_stable(pkg, orc, ctx):
    rng = np.random.default_rng(1)
    L = 70
    S = rng.integers(0, 6, size=(L, L)).astype(np.float64)   # many exact ties
    S = np.triu(S, 1)
    S = S + S.T
    for ms in (1, 4, 5, 69, 70, 200):
        assert pkg.compute_ranking(S, ms, ctx=ctx) == orc.compute_ranking(S, ms)
    S[3, 9] = S[9, 3] = -0.0
    S[4, 9] = S[9, 4] = 0.0
    assert pkg.compute_ranking(-S, 1, ctx=ctx) == orc.compute_ranking(-S, 1)
    Sc = pkg.correct_APC(S, ctx=ctx)
    assert normwise(Sc, orc.correct_APC(S)) <= 1e-14


def test_ranking_large_sort(pkg, orc, ctx):
    rng = np.random.default_rng(2)
    L = 700  # 241k rows: several global bitonic stages
    S = rng.standard_normal((L, L))
    S = S + S.T
    assert pkg.compute_ranking(S, 5, ctx=ctx) == orc.compute_ranking(S, 5)


# ----------------------------------------------------------------------------- error behaviour
def test_error_behaviour_matches_reference(pkg, ctx, tmp_path):
    fa = golden_path("small.fasta.gz")
    with pytest.raises(ValueError, match="invalid pseudocount value: 1.5"):
        pkg.gDCA(fa, pseudocount=1.5, ctx=ctx)
    with pytest.raises(ValueError, match="invalid θ value"):
        pkg.gDCA(fa, theta=2.0, ctx=ctx)
    with pytest.raises(ValueError, match="invalid score value: foo"):
        pkg.gDCA(fa, score="foo", ctx=ctx)
    with pytest.raises(ValueError, match="invalid min_separation value: 0"):
        pkg.gDCA(fa, min_separation=0, ctx=ctx)
    with pytest.raises(ValueError, match="cannot open file"):
        pkg.gDCA(str(tmp_path / "nope.fasta"), ctx=ctx)
    Z = np.full((10, 8), 1, dtype=np.int8)
    Z[0, 0] = 40
    with pytest.raises(pkg.GdcaError, match="parameter q=40 is too big"):
        pkg.gdca_from_alignment(Z, ctx=ctx)
    # the C ABI re-checks ranges itself
    from gaussdca_jl_b200._lib import RANK_DTYPE, ptr
    Zs = np.ones((4, 10), dtype=np.int8)
    R = np.empty(15, dtype=RANK_DTYPE)
    assert ctx.lib.gdca_run(ctx.h, ptr(Zs), 10, 4, -1.0, 2.0, 0, 5, ptr(R), 15, None) == 1
    assert ctx.lib.gdca_run(ctx.h, ptr(Zs), 10, 4, -1.0, 0.5, 0, 5, ptr(R), 14, None) == 1
    assert b"R_len" in ctx.lib.gdca_last_error(ctx.h)


def test_pseudocount_zero_not_spd_end_to_end(pkg, ctx):
    with pytest.raises(pkg.PosDefException):
        pkg.gDCA(golden_path("small.fasta.gz"), pseudocount=0.0, ctx=ctx)
    # and the context is still usable afterwards
    R = pkg.gDCA(golden_path("small.fasta.gz"), ctx=ctx)
    assert R[0][:2] == (11, 35)


# ----------------------------------------------------------------------------- bigger shapes, properties
def test_config_B_slice_end_to_end_vs_oracle(pkg, orc, ctx):
    """BASELINE config B geometry (L=200) at an M the oracle finishes in seconds."""
    Z = orc.synth_alignment(200, 6000, seed=20140321)
    st = {}
    Ro = orc.gdca_from_Z(Z, stages=st)
    R, s = pkg.gdca_from_alignment(Z, ctx=ctx, return_stats=True)
    assert s["thresh"] == st["thresh"] and s["theta"] == st["theta"] and s["meff"] == st["Meff"]
    assert_rank_equal_tie_aware(R, Ro)
    top = [(i, j) for i, j, _ in R[:200]]
    topo = [(i, j) for i, j, _ in Ro[:200]]
    assert top == topo


def test_properties_at_full_config_B(pkg, ctx):
    """L=200, M=50k (BASELINE configs[1]): size-independent properties instead of an oracle run."""
    import ctypes
    from gaussdca_jl_b200._lib import ptr
    L, M = 200, 50000
    Z = np.empty((M, L), dtype=np.int8)
    ctx.check(ctx.lib.gdca_synth_alignment(ctx.h, ptr(Z), L, M, 20140321))
    w = pkg.compute_weights(Z, "auto", ctx=ctx, full=True)
    counts = w["counts"]
    # (1) permutation equivariance of the integer counts
    perm = np.random.default_rng(0).permutation(M)
    w2 = pkg.compute_weights(np.ascontiguousarray(Z[perm]), "auto", ctx=ctx, full=True)
    assert w2["ident_sum"] == w["ident_sum"] and w2["thresh"] == w["thresh"]
    assert np.array_equal(w2["counts"], counts[perm])
    assert w2["Meff"] == w["Meff"]                     # correctly rounded -> order independent
    # (2) sum of (count-1) is even: every neighbour pair credits both members
    assert int((counts.astype(np.int64) - 1).sum()) % 2 == 0
    # (3) appending an exact duplicate of sequence 0 raises count[0] and its neighbours' counts by one
    Z3 = np.concatenate([Z, Z[:1]])
    w3 = pkg.compute_weights(Z3, w["thresh"] / L + 1e-9, ctx=ctx, full=True)
    wfix = pkg.compute_weights(Z, w["thresh"] / L + 1e-9, ctx=ctx, full=True)
    assert w3["thresh"] == w["thresh"]
    d = w3["counts"][:M].astype(np.int64) - wfix["counts"]
    assert d[0] == 1 and set(np.unique(d)) <= {0, 1} and d.sum() == wfix["counts"][0]
    assert w3["counts"][M] == w3["counts"][0]
    # (4) end to end: covariance inverse really is an inverse, ranking sorted and complete
    R, st = pkg.gdca_from_alignment(Z, ctx=ctx, return_stats=True, as_array=True)
    assert len(R) == (L - 5) * (L - 4) // 2
    assert np.all(np.diff(R["score"]) <= 0) and np.all(R["j"] - R["i"] >= 5)
    C, Pi, q = pkg.compute_covariance(Z, w["W"], w["Meff"], 0.8, ctx=ctx)
    mJ = pkg.inverse(C, ctx=ctx)
    assert np.max(np.abs(mJ @ C - np.eye(C.shape[0]))) < 1e-9


# ----------------------------------------------------------------------------- more edges
def test_synthetic_generator_bytes_match_oracle(pkg, orc, ctx):
    from gaussdca_jl_b200._lib import ptr
    for L, M in [(37, 230), (200, 1000), (5, 49)]:
        Z = np.empty((M, L), dtype=np.int8)
        ctx.check(ctx.lib.gdca_synth_alignment(ctx.h, ptr(Z), L, M, 20140321))
        assert np.array_equal(Z, orc.synth_alignment(L, M, 20140321))


def test_q31_thirty_states(pkg, orc, ctx):
    """Largest alphabet the reference allows (q = 31, src/GaussDCA.jl:26): 5 planes, s = 30 blocks."""
    rng = np.random.default_rng(31)
    Z = rng.integers(1, 32, size=(600, 9), dtype=np.int8)
    Z[0, 0] = 31
    for score in ("frob", "DI"):
        st = {}
        Ro = orc.gdca_from_Z(Z, 0.5, "auto", score, 2, stages=st)
        R, s = pkg.gdca_from_alignment(Z, 0.5, "auto", score, 2, ctx=ctx, return_stats=True)
        assert s["q"] == 31 and s["n"] == 270 and s["thresh"] == st["thresh"] and s["meff"] == st["Meff"]
        assert_rank_equal_tie_aware(R, Ro)


def test_degenerate_shapes(pkg, orc, ctx):
    # L <= min_separation: empty ranking, like the reference's zero-length Vector (src/GaussDCA.jl:90)
    Z = orc.synth_alignment(5, 300, seed=1)
    assert pkg.gdca_from_alignment(Z, ctx=ctx) == []
    assert pkg.gdca_from_alignment(Z, min_separation=4, ctx=ctx)[0][:2] == (1, 5)
    # a single sequence: theta must be given, W = 1, C comes from the pseudocount alone
    Z1 = orc.synth_alignment(12, 1, seed=2)
    Z1[0, 0] = 21
    Ro = orc.gdca_from_Z(Z1, 0.8, 0.2, "frob", 3)
    R = pkg.gdca_from_alignment(Z1, 0.8, 0.2, "frob", 3, ctx=ctx)
    assert_rank_equal_tie_aware(R, Ro)
    with pytest.raises(ValueError, match="at least 2 sequences"):
        pkg.gdca_from_alignment(Z1, ctx=ctx)
    # all sequences identical: every pair is a neighbour, Meff = 1
    Zs = np.repeat(orc.synth_alignment(20, 1, seed=3), 500, axis=0)
    Zs[:, 0] = 21
    w = pkg.compute_weights(Zs, "auto", ctx=ctx, full=True)
    assert w["theta"] == orc.compute_theta(Zs) and np.all(w["counts"] == 500) and w["Meff"] == 1.0


def test_resident_run_equals_host_run_and_context_reuse(pkg, orc, ctx):
    """gdca_run_resident (Z already in HBM) == gdca_run (host Z); one context across shapes and scores."""
    import ctypes
    import torch
    from gaussdca_jl_b200 import _lib as glib
    for (L, M, score) in [(64, 5000, "frob"), (33, 700, "DI"), (128, 2000, "frob")]:
        Z = orc.synth_alignment(L, M, seed=L)
        R_host = pkg.gdca_from_alignment(Z, score=score, ctx=ctx, as_array=True)
        Zd = torch.from_numpy(Z).cuda()
        n_out = int(ctx.lib.gdca_ranking_length(L, 5))
        R = np.empty(n_out, dtype=glib.RANK_DTYPE)
        st = glib.Stats()
        ctx.check(ctx.lib.gdca_run_resident(ctx.h, ctypes.c_void_p(Zd.data_ptr()), L, M, -1.0, 0.8,
                                            glib.SCORE_CODES[score], 5, glib.ptr(R), n_out, ctypes.byref(st)))
        assert np.array_equal(R, R_host)          # same kernels, same order: bit-identical
        assert st.theta_passes == 1 and st.ms_total > 0


def test_site_reordering_keeps_counts_exact_on_pfam_like_columns(pkg, orc, ctx):
    """Conserved and gap-heavy columns are packed last (earlier early exit); counts must not change."""
    rng = np.random.default_rng(9)
    L, M = 150, 4000
    Z = orc.synth_alignment(L, M, seed=12)
    cons = rng.choice(L, size=60, replace=False)          # 40 % of the columns conserved / gappy
    for c in cons[:30]:
        Z[:, c] = np.where(rng.random(M) < 0.97, Z[0, c], Z[:, c])
    for c in cons[30:]:
        Z[:, c] = np.where(rng.random(M) < 0.9, 21, Z[:, c])
    Z[0, 0] = 21
    for theta in ("auto", 0.2, 0.45):
        tho = orc.compute_theta(Z) if theta == "auto" else theta
        counts, W, Meff, thresh = orc.compute_weights(Z, tho)
        w = pkg.compute_weights(Z, theta, ctx=ctx, full=True)
        assert w["theta"] == tho and w["thresh"] == thresh
        assert np.array_equal(w["counts"], counts) and w["Meff"] == Meff


# ----------------------------------------------------------------------------- DCAUtils-shaped staged pieces
@pytest.mark.parametrize("L,M,theta", [(53, 300, "auto"), (40, 700, 0.3), (17, 90, 0.0)])
def test_weighted_frequencies_pseudocount_and_C_pieces(pkg, orc, ctx, L, M, theta):
    """compute_weighted_frequencies / add_pseudocount / compute_C one by one (src/GaussDCA.jl:28-32) against the oracle,
    and against the fused covariance stage: the pieces compose to the same C."""
    Z = orc.synth_alignment(L, M, seed=5 + L + M)
    q = int(Z.max())
    Pi_o, Pij_o, Meff_o, W_o, info = orc.compute_weighted_frequencies(Z, q, theta)
    Pi_t, Pij_t, Meff, W = pkg.compute_weighted_frequencies(Z, q, theta, ctx=ctx)
    assert Meff == Meff_o and np.array_equal(W, W_o)
    assert np.array_equal(Pij_t, Pij_t.T)
    assert normwise(Pi_t, Pi_o) <= 1e-14 and normwise(Pij_t, Pij_o) <= 1e-14
    for pc in (0.8, 0.2, 0.0, 1.0):
        Pi_po, Pij_po = orc.add_pseudocount(Pi_o, Pij_o, pc, q)
        Pi_p, Pij_p = pkg.add_pseudocount(Pi_o, Pij_o, pc, q, ctx=ctx)
        assert normwise(Pi_p, Pi_po) <= 1e-15 and normwise(Pij_p, Pij_po) <= 1e-15
        C_o = orc.compute_C(Pi_po, Pij_po)
        C_p = pkg.compute_C(Pi_po, Pij_po, ctx=ctx)
        assert normwise(C_p, C_o) <= 1e-15
        # the pieces chained on the GPU == the fused stage
        C_chain = pkg.compute_C(*pkg.add_pseudocount(Pi_t, Pij_t, pc, q, ctx=ctx), ctx=ctx)
        C_fused, _, qf = pkg.compute_covariance(Z, W, Meff, pc, ctx=ctx)
        assert qf == q
        assert normwise(C_chain, C_fused) <= 1e-13
    with pytest.raises(Exception):
        pkg.add_pseudocount(Pi_o, Pij_o, 1.5, q, ctx=ctx)


# ----------------------------------------------------------------------------- round-2 robustness fixes
def test_apc_on_a_non_symmetric_matrix(pkg, orc, ctx):
    """correct_APC uses row sums x column sums (src/GaussDCA.jl:80-84); the staged entry point takes any S."""
    rng = np.random.default_rng(4)
    S = rng.random((37, 37))
    np.fill_diagonal(S, 0.0)
    assert normwise(pkg.correct_APC(S, ctx=ctx), orc.correct_APC(S)) <= 1e-14
    Ss = S + S.T
    out = pkg.correct_APC(Ss, ctx=ctx)
    assert np.array_equal(out, out.T)     # symmetric in -> symmetric out, bit for bit


def test_ranking_nan_sorts_first_like_julia_isless(pkg, orc, ctx):
    """Julia's sort!(rev=true) uses isless: every NaN (either sign bit) is above +Inf; ties keep enumeration order."""
    L = 12
    S = np.random.default_rng(6).standard_normal((L, L))
    S = S + S.T
    S[5, 1] = S[1, 5] = np.nan
    S[9, 2] = S[2, 9] = -np.nan
    S[8, 3] = S[3, 8] = np.inf
    R = pkg.compute_ranking(S, 1, ctx=ctx)
    Ro = orc.compute_ranking(S, 1)
    assert [(i, j) for i, j, _ in R] == [(i, j) for i, j, _ in Ro]
    assert [(i, j) for i, j, _ in R[:3]] == [(2, 6), (3, 10), (4, 9)]


def test_residue_code_below_one_is_rejected(pkg, ctx):
    Z = np.ones((20, 9), dtype=np.int8)
    Z[3, 4] = -5
    with pytest.raises(ValueError, match="codes must be >= 1"):
        pkg.gdca_from_alignment(Z, ctx=ctx)
    Z[3, 4] = 0
    with pytest.raises(ValueError, match="codes must be >= 1"):
        pkg.compute_weights(Z, 0.2, ctx=ctx)


def test_resident_run_on_an_unaligned_device_view_and_state_drop(pkg, orc, ctx):
    """gdca_run_resident takes any device pointer (odd byte offsets included) and lets go of it when it returns."""
    import ctypes
    import torch
    from gaussdca_jl_b200 import _lib as glib
    L, M = 45, 900
    Z = orc.synth_alignment(L, M, seed=8)
    R_host = pkg.gdca_from_alignment(Z, ctx=ctx, as_array=True)
    buf = torch.zeros(L * M + 64, dtype=torch.int8, device="cuda")
    for off in (1, 7, 16, 33):
        buf[off:off + L * M] = torch.from_numpy(Z.reshape(-1)).cuda()
        n_out = int(ctx.lib.gdca_ranking_length(L, 5))
        R = np.empty(n_out, dtype=glib.RANK_DTYPE)
        ctx.check(ctx.lib.gdca_run_resident(ctx.h, ctypes.c_void_p(buf.data_ptr() + off), L, M, -1.0, 0.8, 0, 5, glib.ptr(R),
                                            n_out, None))
        assert np.array_equal(R, R_host)
    # the borrowed pointer is gone: a device-resident stage now reports a state error instead of reading freed memory
    assert ctx.lib.gdca_dev_covariance(ctx.h, 0.8) == glib.GDCA_ERR_STATE
    # and a staged host-buffer call that changes L invalidates the alignment of an earlier gdca_dev_load
    ctx.check(ctx.lib.gdca_dev_load(ctx.h, glib.ptr(Z), L, M))
    pkg.correct_APC(np.ones((7, 7)), ctx=ctx)
    assert ctx.lib.gdca_dev_pair_pass(ctx.h, 1, 10) == glib.GDCA_ERR_STATE


def test_not_spd_chain_stops_early(pkg, ctx):
    """A failed pivot in the first diagonal block: the launches queued behind it return at once (the reference throws at the
    pivot, src/GaussDCA.jl:34) -- the call still reports LAPACK's info."""
    import time
    n = 4000
    rng = np.random.default_rng(1)
    A = rng.standard_normal((n, n + 5))
    C = A @ A.T / n + np.eye(n)
    pkg.inverse(C, ctx=ctx)                     # warm-up (allocations)
    t0 = time.perf_counter(); pkg.inverse(C, ctx=ctx); t_ok = time.perf_counter() - t0
    C[2, 2] = -1.0
    t0 = time.perf_counter()
    with pytest.raises(pkg.PosDefException) as ei:
        pkg.inverse(C, ctx=ctx)
    t_bad = time.perf_counter() - t0
    assert ei.value.info == 3
    assert t_bad < t_ok                          # both include the same H2D of C; the failed run skips D2H and the flop
