"""Parity at the HEADLINE sizes (BASELINE.json configs[2] and configs[3]: L=500, M=200 000), not at a slice.

The north star's own acceptance test: "gDCA end-to-end on a synthetic L=500, M=200k alignment with weights bit-exact,
the top-L ranking identical to the reference".  The oracle (oracle/fullsize.py = oracle.gdca_oracle stage by stage on
all host cores; reference pipeline src/GaussDCA.jl:24-46, comparison contract test/runtests.jl:41-50) is run FOR REAL
on the full alignment: two O(M^2 L) packed pair sweeps (~100 s on 16 threads), scatter-add frequencies, LAPACK
dpotrf+dpotri at n = 10 000, FN / DI, APC, stable ranking.  One oracle run per box is shared by all tests here and by
bench.py through a cache of the oracle's own outputs.

Tolerances: counts / W / theta / thresh / Meff bit-exact; C, mJ, raw S, APC S normwise max|d|/max|.| <= 1e-9; ranking
identical up to mathematical ties, top-L (L = 500 pairs) identical as a list.
"""
import numpy as np
import pytest

from test_gpu_parity import TOL, assert_rank_equal_tie_aware, normwise

pytestmark = pytest.mark.gpu

L, M, SEED = 500, 200_000, 20140321
CONFIGS = {"C": ("frob", 0.8), "D": ("DI", 0.2)}   # BASELINE.json configs[2], configs[3]


@pytest.fixture(scope="module")
def full():
    import __graft_entry__ as g
    g.load_oracle().build()
    from oracle import fullsize
    return fullsize


@pytest.fixture(scope="module")
def Z(orc):
    return orc.synth_alignment(L, M, SEED)


@pytest.fixture(scope="module")
def gpu_weights(pkg, ctx, Z):
    return pkg.compute_weights(Z, "auto", ctx=ctx, full=True)


def test_config_C_weights_bitexact_vs_full_oracle_sweep(full, Z, gpu_weights):
    """compute_theta + compute_weights at L=500, M=200k: every integer equal, W and Meff the same doubles."""
    o = full.weights_full(L, M, SEED, Z=Z)
    w = gpu_weights
    assert w["ident_sum"] == int(o["ident_sum"])
    assert w["theta"] == o["theta"] and w["thresh"] == o["thresh"]
    assert int(np.max(np.abs(w["counts"].astype(np.int64) - o["counts"]))) == 0
    assert np.array_equal(w["W"], 1.0 / o["counts"].astype(np.float64))
    assert w["Meff"] == o["Meff"]
    assert w["passes"] == 1          # theta=:auto costs no pair sweep on the GPU (site histograms), the oracle does two


@pytest.mark.parametrize("cfg", ["C", "D"])
def test_config_pipeline_vs_full_oracle(pkg, ctx, full, Z, gpu_weights, cfg):
    """configs[2] (:frob, pc 0.8) and configs[3] (:DI, pc 0.2) end to end and stage by stage against the oracle."""
    score, pc = CONFIGS[cfg]
    o = full.pipeline_full(L, M, score, pc, SEED, Z=Z, keep_big=True)
    w = gpu_weights
    # covariance (frequencies + pseudocount + C fused)
    C, Pi, q = pkg.compute_covariance(Z, w["W"], w["Meff"], pc, ctx=ctx)
    assert q == o["q"] == 21
    assert np.array_equal(C, C.T)
    if "C" in o:
        assert normwise(C, o["C"]) <= 1e-12
    assert float(np.max(np.abs(C[::full.SAMPLE_STRIDE] - o["C_rows"]))) / o["C_absmax"] <= 1e-12
    # inverse
    mJ = pkg.inverse(C, ctx=ctx)
    assert np.array_equal(mJ, mJ.T)
    if "mJ" in o:
        assert normwise(mJ, o["mJ"]) <= TOL
    assert float(np.max(np.abs(mJ[::full.SAMPLE_STRIDE] - o["mJ_rows"]))) / o["mJ_absmax"] <= TOL
    # block scores from the GPU's own inverse, then APC
    S = pkg.compute_DI_gauss(mJ, C, q, ctx=ctx) if score == "DI" else pkg.compute_FN(mJ, q, ctx=ctx)
    del C, mJ
    assert normwise(S, o["S_raw"]) <= TOL
    assert normwise(pkg.correct_APC(S, ctx=ctx), o["S"]) <= TOL
    # the fused call a user makes (gdca_run): ranking tie-aware identical, top-L identical as a list
    R, st = pkg.gdca_from_alignment(Z, pc, "auto", score, 5, ctx=ctx, return_stats=True)
    assert st["thresh"] == o["thresh"] and st["theta"] == o["theta"] and st["meff"] == o["Meff"]
    Ro = [(int(i), int(j), float(x)) for i, j, x in o["R"].tolist()]
    assert len(R) == len(Ro) == (L - 5) * (L - 4) // 2
    assert_rank_equal_tie_aware(R, Ro)
    assert [(i, j) for i, j, _ in R[:L]] == [(i, j) for i, j, _ in Ro[:L]]


def test_config_C_shuffled_sequence_order(pkg, ctx, full, Z, gpu_weights):
    """The same multiset of sequences in random order: family members no longer sit on the lines k = l (mod 4000), so the
    prefilter flags far more blocks.  Counts are the permuted counts bit for bit, Meff the same double,
    and the ranking equals the oracle's up to summation-order noise (<= 1e-9 normwise, tie-aware)."""
    o = full.pipeline_full(L, M, "frob", 0.8, SEED, Z=Z)
    perm = np.random.default_rng(SEED).permutation(M)
    Zs = np.ascontiguousarray(Z[perm])
    w = pkg.compute_weights(Zs, "auto", ctx=ctx, full=True)
    assert w["ident_sum"] == int(o["ident_sum"]) and w["theta"] == o["theta"] and w["thresh"] == o["thresh"]
    assert np.array_equal(w["counts"], o["counts"][perm])
    assert w["Meff"] == o["Meff"]
    R = pkg.gdca_from_alignment(Zs, 0.8, "auto", "frob", 5, ctx=ctx)
    Ro = [(int(i), int(j), float(x)) for i, j, x in o["R"].tolist()]
    assert_rank_equal_tie_aware(R, Ro)
    assert [(i, j) for i, j, _ in R[:L]] == [(i, j) for i, j, _ in Ro[:L]]
