"""CPU: pins the oracle against every golden vector the reference's tests hold (test/runtests.jl:52-86)."""
import json
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_CASES, golden_path, printed_todict, read_golden


@pytest.mark.parametrize("name,fa,kw", GOLDEN_CASES, ids=[c[0] for c in GOLDEN_CASES])
def test_oracle_reproduces_golden(orc, name, fa, kw):
    """Same comparison as the reference: dict keyed by (i,j), key sets equal, values equal at the 7 digits
    of the '%e' print (compare_results, test/runtests.jl:41-50)."""
    R = orc.gDCA(golden_path(fa), **kw)
    got = printed_todict(orc.format_rank(R))
    want = read_golden(name)
    assert sorted(got) == sorted(want)
    assert all(math.isclose(got[k], want[k], rel_tol=1.5e-8, abs_tol=0.0) for k in want)
    xs = [x for _, _, x in R]
    assert all(xs[t] >= xs[t + 1] for t in range(len(xs) - 1))


def test_oracle_fallback_path_matches_packed_path(orc):
    """test3 of the reference (DCAUTILS_FORCE_FALLBACK, test/runtests.jl:78-86): the un-packed pair
    sweep gives the same answer as the packed one."""
    kw = dict(pseudocount=0.2, score="DI", remove_dups=True)
    a = orc.gDCA(golden_path("small.fasta.gz"), packed=True, **kw)
    b = orc.gDCA(golden_path("small.fasta.gz"), packed=False, **kw)
    assert a == b
    want = read_golden("small.DIRout")
    got = printed_todict(orc.format_rank(b))
    assert sorted(got) == sorted(want) and all(math.isclose(got[k], want[k], rel_tol=1.5e-8) for k in want)


def test_stage_anchors(orc):
    """Secondary anchors: theta, thresh, Meff, count histogram, entries of C and mJ."""
    anchors = json.load(open(os.path.join(GOLDEN, "stage_anchors.json")))
    a = anchors["small_default"]
    st = {}
    R = orc.gDCA(golden_path("small.fasta.gz"), stages=st)
    assert st["theta"] == a["theta"] and st["thresh"] == a["thresh"] == 19 and st["Meff"] == a["Meff"] == 92.0
    vals, h = np.unique(st["counts"], return_counts=True)
    assert {str(int(v)): int(c) for v, c in zip(vals, h)} == a["count_hist"] == {"1": 82, "2": 14, "3": 6, "4": 4}
    assert math.isclose(st["C"][0, 0], 0.04262630256891667, rel_tol=1e-13)
    assert math.isclose(np.trace(st["C"]), 47.47735212011711, rel_tol=1e-13)
    assert math.isclose(st["mJ"][0, 0], 42.78127152455032, rel_tol=1e-12)
    assert R[0][:2] == (11, 35) and math.isclose(R[0][2], 3.6494745366789094, rel_tol=1e-12)
    b = anchors["large_DI_dedup"]
    assert b["M"] == 94 and b["L"] == 400 and b["thresh"] == 147 and math.isclose(b["Meff"], 25.8138802992673, rel_tol=1e-14)


def test_oracle_pieces_small_cases(orc):
    rng = np.random.default_rng(0)
    Z = rng.integers(1, 22, size=(40, 13), dtype=np.int8)
    # packed and byte sweeps agree; brute force agrees
    H = (Z[:, None, :] != Z[None, :, :]).sum(-1)
    iu = np.triu_indices(40, 1)
    assert orc.ident_sum(Z, True) == orc.ident_sum(Z, False) == int((13 - H[iu]).sum())
    for th in (0.1, 0.5, 1.0):
        thresh = int(math.floor(th * 13))
        want = (H < thresh).sum(1).astype(np.int32)       # includes self (H=0 < thresh when thresh >= 1)
        if thresh == 0:
            want = np.ones(40, dtype=np.int32)
        for packed in (True, False):
            counts, W, Meff, t = orc.compute_weights(Z, th, packed)
            assert t == thresh and np.array_equal(counts, want) and np.array_equal(W, 1.0 / want)
    # theta == 0: no sweep, unit weights
    c, W, Meff, t = orc.compute_weights(Z, 0.0)
    assert Meff == 40.0 and np.all(W == 1.0)
    # frequencies against the dense one-hot contraction
    q = 21
    W = rng.random(40) + 0.1
    Meff = float(W.sum())
    Pi, Pij = orc.compute_freqs(Z, q, W, Meff)
    X = np.zeros((40, 13 * 20))
    for k in range(40):
        for i in range(13):
            if Z[k, i] < q:
                X[k, i * 20 + int(Z[k, i]) - 1] = 1
    assert np.allclose(Pij, X.T @ (W[:, None] * X) / Meff, rtol=1e-13, atol=1e-15)
    assert np.allclose(Pi, X.T @ W / Meff, rtol=1e-13)
    # empty / degenerate theta: all sequences different everywhere -> meanfracid 0 -> theta capped at 0.5
    Zd = np.arange(1, 5, dtype=np.int8).reshape(4, 1)
    assert orc.compute_theta(Zd) == 0.5


def test_meff_is_correctly_rounded(orc):
    from fractions import Fraction
    counts = np.array([3] * 7 + [7] * 5 + [1] * 11 + [13], dtype=np.int32)
    exact = Fraction(7, 3) + Fraction(5, 7) + 11 + Fraction(1, 13)
    assert orc.meff_from_counts(counts) == float(exact)


def test_synth_generator_matches_numpy_restatement(orc):
    """The generator is specified, not just implemented: a vectorised numpy restatement gives the same bytes."""
    L, M, seed = 37, 230, 20140321
    Z = orc.synth_alignment(L, M, seed)
    U = np.uint64

    def sm(x):
        x = (x + U(0x9E3779B97F4A7C15)).astype(U)
        x = ((x ^ (x >> U(30))) * U(0xBF58476D1CE4E5B9)).astype(U)
        x = ((x ^ (x >> U(27))) * U(0x94D049BB133111EB)).astype(U)
        return x ^ (x >> U(31))

    def draw(t, a, b):
        with np.errstate(over="ignore"):
            return sm(sm(sm(np.full_like(a, U(seed) ^ (U(t) * U(0xD1B54A32D192ED03)))) + a) + b)

    def u01(r):
        return (r >> U(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    with np.errstate(over="ignore"):
        k = np.arange(M, dtype=U)[:, None] * np.ones((1, L), dtype=U)
        i = np.ones((M, 1), dtype=U) * np.arange(L, dtype=U)[None, :]
        K = U(max(M // 50, 1))
        anc = k % K
        mu = 0.05 + 0.60 * u01(draw(3, k, np.zeros_like(k)))
        ra = draw(1, anc, i)
        v = np.where(u01(ra) < 0.10, 21, 1 + (sm(ra) % U(20)).astype(np.int64))
        rm = draw(2, k, i)
        v = np.where(u01(rm) < mu, 1 + (sm(rm) % U(21)).astype(np.int64), v)
    v[0, 0] = 21
    assert np.array_equal(Z, v.astype(np.int8))
    assert Z.max() == 21 and Z.min() >= 1
