"""The CPU oracle run for real at a full BASELINE.json config (no sampling, no extrapolation), with stage timers and a
per-box cache.  TEST INFRASTRUCTURE ONLY: imported by tests/test_gpu_fullsize.py and by the cpu_baseline /
--impl reference / parity legs of bench.py, never by the product.

What runs is exactly oracle.gdca_oracle.gdca_from_Z (reference src/GaussDCA.jl:24-46, DCAUtils call sites :28,:30,:37,:39)
stage by stage on all host threads: compress_Z, compute_theta (pair sweep 1), compute_weights (pair sweep 2),
frequencies, add_pseudocount, compute_C, LAPACK dpotrf+dpotri, FN / DI, APC, ranking.  At config C (L=500, M=200k) that
is ~100 s on 16 threads, almost all of it in the two pair sweeps, so the small results are cached under
$GDCA_ORACLE_CACHE (default /tmp/gdca_oracle_cache) keyed by the config: the reference arm of bench.py, the parity leg of
the repo's arm and the full-size tests then share ONE oracle run per box.  The cache holds only results of the oracle
itself (never anything the GPU produced), together with the seconds each stage took when it was computed on this box.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

from . import gdca_oracle as orc

CACHE_DIR = os.environ.get("GDCA_ORACLE_CACHE", "/tmp/gdca_oracle_cache")
SAMPLE_STRIDE = 53  # rows of C / mJ kept in the cache (every 53rd row): a cache hit can still check the big matrices


def set_all_threads():
    """Use every host core: launchers such as torchrun export OMP_NUM_THREADS=1."""
    n = os.cpu_count() or 1
    orc.lib().oracle_set_threads(n)
    return n


class _blas_threads:
    """SciPy's OpenBLAS reads OMP_NUM_THREADS at load time; raise its pool to all cores around the LAPACK calls."""

    def __init__(self, n):
        self.n, self.cm = n, None

    def __enter__(self):
        try:
            from threadpoolctl import threadpool_limits
            self.cm = threadpool_limits(limits=self.n)
            self.cm.__enter__()
        except Exception:
            self.cm = None
        return self

    def __exit__(self, *a):
        if self.cm is not None:
            self.cm.__exit__(*a)


def _key(L, M, seed, extra=""):
    return f"L{L}_M{M}_s{seed}{extra}"


def _load(path):
    try:
        with np.load(path, allow_pickle=False) as z:
            return {k: z[k] for k in z.files}
    except Exception:
        return None


def _save(path, d):
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = f"{path}.{os.getpid()}.tmp.npz"
        np.savez(tmp, **d)
        os.replace(tmp, path)
    except Exception:
        pass


def weights_full(L, M, seed=20140321, Z=None, use_cache=True):
    """compute_theta + compute_weights on the whole alignment (two O(M^2 L) pair sweeps, packed 5-bit path).
    -> dict(counts, theta, thresh, Meff, ident_sum, threads, t_pack, t_theta, t_counts, cached)"""
    path = os.path.join(CACHE_DIR, _key(L, M, seed) + "_weights.npz")
    if use_cache:
        d = _load(path)
        if d is not None:
            out = {k: (v if v.ndim else v.item()) for k, v in d.items()}
            out["cached"] = True
            return out
    threads = set_all_threads()
    if Z is None:
        Z = orc.synth_alignment(L, M, seed)
    lib = orc.lib()
    t0 = time.perf_counter()
    cZ = orc.compress_Z(Z)
    t_pack = time.perf_counter() - t0
    t0 = time.perf_counter()
    ident = int(lib.oracle_ident_sum_packed(orc._ptr(cZ), L, M))
    theta = orc.theta_from_ident_sum(ident, L, M)
    t_theta = time.perf_counter() - t0
    thresh = int(np.floor(theta * L))
    counts = np.empty(M, dtype=np.int32)
    t0 = time.perf_counter()
    lib.oracle_neighbour_counts_packed(orc._ptr(cZ), L, M, thresh, orc._ptr(counts))
    Meff = orc.meff_from_counts(counts)
    t_counts = time.perf_counter() - t0
    out = dict(counts=counts, theta=float(theta), thresh=thresh, Meff=float(Meff), ident_sum=np.uint64(ident),
               threads=threads, t_pack=t_pack, t_theta=t_theta, t_counts=t_counts)
    _save(path, out)
    out["ident_sum"] = ident
    out["cached"] = False
    return out


def pipeline_full(L, M, score, pc, seed=20140321, min_separation=5, Z=None, use_cache=True, keep_big=False):
    """The whole oracle pipeline at (L, M).  -> dict with the weights_full entries plus
    S_raw, S (APC), R (structured i/j/score), C_rows / mJ_rows (every SAMPLE_STRIDE-th row), C_absmax, mJ_absmax,
    t_freqs, t_pc_C, t_inv, t_score, t_apc_rank, t_total, and (keep_big, fresh runs only) the full C and mJ."""
    w = weights_full(L, M, seed, Z=Z, use_cache=use_cache)
    path = os.path.join(CACHE_DIR, _key(L, M, seed, f"_{score}_pc{pc}_ms{min_separation}") + "_pipe.npz")
    d = _load(path) if (use_cache and not keep_big) else None
    if d is None:
        threads = set_all_threads()
        if Z is None:
            Z = orc.synth_alignment(L, M, seed)
        q = int(Z.max())
        W = 1.0 / w["counts"].astype(np.float64)
        t0 = time.perf_counter()
        Pi_true, Pij_true = orc.compute_freqs(Z, q, W, w["Meff"])
        t_freqs = time.perf_counter() - t0
        t0 = time.perf_counter()
        Pi, Pij = orc.add_pseudocount(Pi_true, Pij_true, float(pc), q)
        del Pij_true
        C = orc.compute_C(Pi, Pij)
        del Pij
        t_pc_C = time.perf_counter() - t0
        t0 = time.perf_counter()
        with _blas_threads(threads):
            mJ = orc.inv_cholesky(C)
        t_inv = time.perf_counter() - t0
        t0 = time.perf_counter()
        with _blas_threads(threads):
            S_raw = orc.compute_DI_gauss(mJ, C, q) if score == "DI" else orc.compute_FN(mJ, q)
        t_score = time.perf_counter() - t0
        t0 = time.perf_counter()
        S = orc.correct_APC(S_raw)
        Rl = orc.compute_ranking(S, min_separation)
        t_apc_rank = time.perf_counter() - t0
        R = np.array(Rl, dtype=[("i", np.int64), ("j", np.int64), ("score", np.float64)])
        d = dict(q=q, S_raw=S_raw, S=S, R=R, C_rows=C[::SAMPLE_STRIDE].copy(), mJ_rows=mJ[::SAMPLE_STRIDE].copy(),
                 C_absmax=float(np.abs(C).max()), mJ_absmax=float(np.abs(mJ).max()), t_freqs=t_freqs, t_pc_C=t_pc_C,
                 t_inv=t_inv, t_score=t_score, t_apc_rank=t_apc_rank)
        _save(path, d)
        d["cached"] = False
        if keep_big:
            d["C"], d["mJ"] = C, mJ
    else:
        d = {k: (v if v.ndim else v.item()) for k, v in d.items()}
        d["cached"] = True
    out = dict(w)
    out.update(d)
    out["weights_cached"] = w["cached"]
    out["t_total"] = (w["t_pack"] + w["t_theta"] + w["t_counts"] + out["t_freqs"] + out["t_pc_C"] + out["t_inv"] +
                      out["t_score"] + out["t_apc_rank"])
    return out


def summary(d):
    """json-able description of a pipeline_full result (bench lines)."""
    keys = ("theta", "thresh", "Meff", "threads", "t_pack", "t_theta", "t_counts", "t_freqs", "t_pc_C", "t_inv", "t_score",
            "t_apc_rank", "t_total", "cached", "weights_cached")
    return json.loads(json.dumps({k: (float(d[k]) if isinstance(d[k], (np.floating, float)) else
                                      int(d[k]) if isinstance(d[k], (np.integer, int)) and not isinstance(d[k], bool) else
                                      bool(d[k])) for k in keys if k in d}))
