"""The frequency / covariance stage on the tensor cores (csrc/covtc.cu): exact co-occurrence counts per weight class with
tcgen05 kind::mxf4, combined in FP64 -- against the oracle (DCAUtils compute_weighted_frequencies, add_pseudocount, compute_C;
call sites src/GaussDCA.jl:28-32) and against the scatter-add engine (csrc/cov.cu) on the same inputs.

Tolerances: Pij_true / C normwise <= 1e-14 (both engines sum the same exact products; only the FP64 summation order differs),
exact symmetry, exact zeros where two different states of ONE site meet; everything downstream (scores, ranking) <= 1e-11."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def normwise(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


@pytest.fixture()
def tc(ctx):
    ctx.set_cov_engine(2)
    yield ctx
    ctx.set_cov_engine(0)


# n = 20 L:  1060 (9 blocks: odd, padded super-row), 800, 340, 2600 (21 blocks), 128 (one block: a lone diagonal super-tile)
@pytest.mark.parametrize("L,M,theta", [(53, 300, "auto"), (40, 3000, 0.3), (17, 90, 0.0), (130, 5000, "auto"), (128, 2500, 0.45),
                                       (7, 1200, "auto")])
def test_weighted_frequencies_on_tensor_cores_vs_oracle_and_scatter_engine(pkg, orc, tc, L, M, theta):
    Z = orc.synth_alignment(L, M, seed=11 + L + M)
    q = int(Z.max())
    Pi_o, Pij_o, Meff_o, W_o, _ = orc.compute_weighted_frequencies(Z, q, theta)
    Pi_t, Pij_t, Meff, W = pkg.compute_weighted_frequencies(Z, q, theta, ctx=tc)
    info = tc.cov_info()
    assert info["engine"] == 2 and info["classes"] >= 1 and info["segments"] >= info["classes"], info
    assert Meff == Meff_o and np.array_equal(W, W_o)
    assert np.array_equal(Pij_t, Pij_t.T)
    assert normwise(Pi_t, Pi_o) <= 1e-14 and normwise(Pij_t, Pij_o) <= 1e-14, (normwise(Pij_t, Pij_o), info)
    # two different states of the same site never co-occur: exact zeros, not rounding noise
    s = q - 1
    for i in range(0, L, max(1, L // 7)):
        blk = Pij_t[i * s:(i + 1) * s, i * s:(i + 1) * s]
        assert np.array_equal(blk, np.diag(np.diag(blk)))
        assert np.allclose(np.diag(blk), Pi_t[i * s:(i + 1) * s], rtol=1e-13, atol=0)
    tc.set_cov_engine(1)
    Pi_s, Pij_s, _, _ = pkg.compute_weighted_frequencies(Z, q, theta, ctx=tc)
    assert tc.cov_info()["engine"] == 1
    tc.set_cov_engine(2)
    assert normwise(Pi_t, Pi_s) <= 1e-15
    assert normwise(Pij_t, Pij_s) <= 1e-14


@pytest.mark.parametrize("L,M,theta,score,pc", [(128, 6000, "auto", "frob", 0.8), (130, 3000, 0.3, "DI", 0.2), (64, 500, 0.0, "frob", 0.5),
                                                (90, 2000, "auto", "DI", 0.95)])
def test_end_to_end_on_tensor_cores_equals_scatter_engine_and_oracle(pkg, orc, tc, L, M, theta, score, pc):
    Z = orc.synth_alignment(L, M, seed=L + M)
    R2, s2 = pkg.gdca_from_alignment(Z, pc, theta, score, 5, ctx=tc, return_stats=True, as_array=True)
    assert tc.cov_info()["engine"] == 2
    tc.set_cov_engine(1)
    R1, s1 = pkg.gdca_from_alignment(Z, pc, theta, score, 5, ctx=tc, return_stats=True, as_array=True)
    tc.set_cov_engine(2)
    assert s1["meff"] == s2["meff"] and s1["thresh"] == s2["thresh"]
    d1 = {(int(i), int(j)): x for i, j, x in R1.tolist()}
    d2 = {(int(i), int(j)): x for i, j, x in R2.tolist()}
    assert sorted(d1) == sorted(d2)
    smax = max(abs(x) for x in d1.values())
    assert max(abs(d1[k] - d2[k]) for k in d1) / smax <= 1e-11
    Ro = orc.gdca_from_Z(Z, pseudocount=pc, theta=theta, score=score, min_separation=5)
    do = {(i, j): x for i, j, x in Ro}
    assert max(abs(do[k] - d2[k]) for k in do) / smax <= 1e-9


def test_goldens_on_tensor_cores(pkg, tc):
    """The reference's own golden files (test/runtests.jl:41-76) with the covariance forced onto the tensor cores."""
    from conftest import GOLDEN_CASES, golden_path, read_golden
    for name, fa, kw in GOLDEN_CASES:
        R = pkg.gDCA(golden_path(fa), ctx=tc, **kw)
        assert tc.cov_info()["engine"] == 2, name
        want = read_golden(name)
        got = {(i, j): x for i, j, x in R}
        assert sorted(got) == sorted(want), name
        for k, v in want.items():
            assert float("%e" % got[k]) == pytest.approx(v, rel=2e-6, abs=0), (name, k)


def test_small_alphabets_and_q31(pkg, orc, tc):
    rng = np.random.default_rng(3)
    for q, L, M in [(20, 33, 800), (31, 21, 900), (3, 70, 600), (2, 150, 400)]:
        Z = rng.integers(1, q + 1, size=(M, L), dtype=np.int8)
        Z[1::3] = Z[0]                      # neighbours: several weight classes
        Z[0, 0] = q
        Pi_o, Pij_o, Meff_o, W_o, _ = orc.compute_weighted_frequencies(Z, q, 0.3)
        Pi_t, Pij_t, Meff, W = pkg.compute_weighted_frequencies(Z, q, 0.3, ctx=tc)
        assert tc.cov_info()["engine"] == 2
        assert Meff == Meff_o and np.array_equal(W, W_o)
        assert np.array_equal(Pij_t, Pij_t.T) and normwise(Pij_t, Pij_o) <= 1e-14, (q, normwise(Pij_t, Pij_o))


def test_degenerate_shapes_on_tensor_cores(pkg, orc, tc):
    """One site, two sequences, a single class, a class per sequence, M far below one k-block: the engine pads, never mixes."""
    rng = np.random.default_rng(5)
    for L, M, theta in [(1, 5, 0.0), (2, 2, 0.3), (3, 40, "auto"), (9, 257, 0.2), (31, 513, 0.9), (5, 1000, 1.0)]:
        Z = rng.integers(1, 22, size=(M, L), dtype=np.int8)
        Z[0, 0] = 21
        q = int(Z.max())
        Pi_o, Pij_o, Meff_o, W_o, _ = orc.compute_weighted_frequencies(Z, q, theta)
        Pi_t, Pij_t, Meff, W = pkg.compute_weighted_frequencies(Z, q, theta, ctx=tc)
        assert tc.cov_info()["engine"] == 2, (L, M)
        assert Meff == Meff_o and np.array_equal(W, W_o)
        assert np.array_equal(Pij_t, Pij_t.T)
        assert normwise(Pi_t, Pi_o) <= 1e-14 and normwise(Pij_t, Pij_o) <= 1e-14, (L, M, theta)


def test_prefilter_and_inversion_report_how_they_ran(pkg, orc, ctx):
    """The default launch modes of this build: CTA pairs with cta_group::2 for the prefilter, the sliced INT8 engine for n >= 2048."""
    Z = orc.synth_alignment(128, 17000, seed=3)        # M >= 16384: prefilter on; n = 2560: sliced inversion
    pkg.gdca_from_alignment(Z, ctx=ctx)
    assert ctx.lib.gdca_dev_tc_filter_launch_mode(ctx.h) == 2
    import ctypes
    mode = ctypes.c_int32()
    ctx.check(ctx.lib.gdca_dev_inverse_info(ctx.h, ctypes.byref(mode), None, None))
    assert mode.value == 1


def test_many_weight_classes_fall_back_to_the_scatter_engine(pkg, orc, tc):
    """More than 512 distinct neighbour counts: engine 2 is a request, the scatter-add engine still answers."""
    L, groups = 24, 560
    rows = []
    rng = np.random.default_rng(9)
    for g in range(1, groups + 1):          # group g: g identical sequences -> count g
        seq = rng.integers(1, 21, size=L, dtype=np.int8)
        rows.extend([seq] * g)
    Z = np.array(rows, dtype=np.int8)
    Z[0, 0] = 21
    Pi_o, Pij_o, Meff_o, W_o, _ = orc.compute_weighted_frequencies(Z, 21, 0.1)
    Pi_t, Pij_t, Meff, W = pkg.compute_weighted_frequencies(Z, 21, 0.1, ctx=tc)
    info = tc.cov_info()
    assert info["engine"] == 1 and info["classes"] > 512, info
    assert Meff == Meff_o and normwise(Pij_t, Pij_o) <= 1e-14


def test_caller_supplied_weights_use_the_scatter_engine(pkg, orc, tc):
    Z = orc.synth_alignment(30, 400, seed=2)
    W = np.random.default_rng(1).random(400)
    C, Pi, q = pkg.compute_covariance(Z, W, float(W.sum()), 0.5, ctx=tc)
    assert tc.cov_info()["engine"] == 1
    Pi_o, Pij_o = orc.compute_freqs(Z, q, W, float(W.sum()))
    C_o = orc.compute_C(*orc.add_pseudocount(Pi_o, Pij_o, 0.5, q))
    assert normwise(C, C_o) <= 1e-13


@pytest.mark.parametrize("members", [2, 3])
def test_group_on_tensor_cores_is_bit_identical_to_single(pkg, orc, ctx, monkeypatch, members):
    monkeypatch.setenv("GDCA_GROUP_ALLOW_SAME_DEVICE", "1")
    ctxg = pkg.Context(devices=[0] * members)
    ctx.set_cov_engine(2)
    ctxg.set_cov_engine(2)
    try:
        for (L, M, theta, score, pc) in [(128, 6000, "auto", "frob", 0.8), (130, 3000, 0.3, "DI", 0.2), (40, 20000, "auto", "frob", 0.8)]:
            Z = orc.synth_alignment(L, M, seed=L + M)
            R1 = pkg.gdca_from_alignment(Z, pc, theta, score, 5, ctx=ctx, as_array=True)
            assert ctx.cov_info()["engine"] == 2
            Rg = pkg.gdca_from_alignment(Z, pc, theta, score, 5, ctx=ctxg, as_array=True)
            assert ctxg.cov_info()["engine"] == 2
            assert np.array_equal(Rg["i"], R1["i"]) and np.array_equal(Rg["j"], R1["j"]), (L, M)
            assert np.array_equal(Rg["score"], R1["score"]), (L, M, float(np.max(np.abs(Rg["score"] - R1["score"]))))
    finally:
        ctx.set_cov_engine(0)
        ctxg.close()
