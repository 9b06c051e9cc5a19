"""CPU oracle for the gDCA hot path.  TEST INFRASTRUCTURE ONLY -- never imported by the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module.  The shipped path is the CUDA library (``gaussdca.jl_b200``),
which has no CPU fallback and fails loudly when its extension is missing.

What is restated, and from where
--------------------------------
The reference (``/root/reference/src/GaussDCA.jl``, 101 lines) is a thin Julia wrapper; most of the
arithmetic lives in the third-party package **DCAUtils 1.x** (uuid e41cd558-3099-4f6e-a65d-5336857e40aa,
``Project.toml:6,12`` -- compat ``"1"``, no Manifest, so no exact version pin) and in Julia's
``LinearAlgebra`` (OpenBLAS LAPACK).  Neither is vendored under ``/root/reference`` and Julia is not
installed in this image, so the reference itself cannot run here.  Each function below cites the
reference line it follows; DCAUtils functions are restated from their published behaviour and cite
the reference's *call site*.

Parity pin
----------
``tests/test_oracle_golden.py`` runs this oracle end to end on the reference's own fixtures
(``test/data/{small,large}.fasta.gz``, committed copies under ``tests/golden/``) and reproduces all four
golden outputs (``small.FNRout.txt``, ``small.DIRout.txt``, ``small.DIRout2.txt``, ``large.DIRout.txt``,
``test/runtests.jl:52-76``) at the 7 significant digits the files carry, with identical key sets.
Intermediates (counts, W, Meff, theta, C, mJ) are **not pinned** by any reference test;
for them this oracle is a validated restatement, not a reference run.
"""
from __future__ import annotations

import ctypes
import gzip
import math
import os
import subprocess
from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgdca_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc + OpenMP).  Building the checker is not using it."""
    src = os.path.join(_HERE, "gdca_oracle_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O3", "-march=x86-64-v2", "-mpopcnt", "-fopenmp", "-shared", "-fPIC", "-o", _SO, src]
        )
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64, i32, u64, dbl = ctypes.c_int64, ctypes.c_int32, ctypes.c_uint64, ctypes.c_double
        p = ctypes.c_void_p
        L.oracle_words_per_seq.restype = i64
        L.oracle_words_per_seq.argtypes = [i64]
        L.oracle_max_threads.restype = ctypes.c_int
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        L.oracle_compress_Z.argtypes = [p, i64, i64, p]
        L.oracle_ident_sum_packed.restype = u64
        L.oracle_ident_sum_packed.argtypes = [p, i64, i64]
        L.oracle_ident_sum_packed_range.restype = u64
        L.oracle_ident_sum_packed_range.argtypes = [p, i64, i64, i64, i64]
        L.oracle_ident_sum_bytes.restype = u64
        L.oracle_ident_sum_bytes.argtypes = [p, i64, i64]
        L.oracle_neighbour_counts_packed.argtypes = [p, i64, i64, i64, p]
        L.oracle_neighbour_counts_packed_range.argtypes = [p, i64, i64, i64, i64, i64, p]
        L.oracle_neighbour_counts_bytes.argtypes = [p, i64, i64, i64, p]
        L.oracle_weighted_freqs.argtypes = [p, i64, i64, i32, p, dbl, p, p]
        L.oracle_weighted_freqs_range.argtypes = [p, i64, i64, i32, p, dbl, i64, i64, p, p]
        L.oracle_synth_alignment.argtypes = [p, i64, i64, u64]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------------------------
# Host I/O: FASTA -> Int8 alignment.  Reference call sites src/GaussDCA.jl:20-23 (DCAUtils
# read_fasta_alignment / remove_duplicate_sequences, un-vendored).
# --------------------------------------------------------------------------------------------
# A..Y -> 1..20 for the 20 standard amino acids, everything else (B J O U X, Z, '-', ...) -> 21.
_ALPHABET = "ACDEFGHIKLMNPQRSTVWY"
_LETTER2NUM = np.full(256, 21, dtype=np.int8)
for _n, _c in enumerate(_ALPHABET):
    _LETTER2NUM[ord(_c)] = _n + 1


def read_fasta_records(path: str):
    opener = gzip.open if path.endswith(".gz") else open
    name, chunks = None, []
    with opener(path, "rt") as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                if name is not None:
                    yield name, "".join(chunks)
                name, chunks = line[1:], []
            else:
                chunks.append(line)
    if name is not None:
        yield name, "".join(chunks)


def read_fasta_alignment(path: str, max_gap_fraction: float) -> np.ndarray:
    """-> Z, int8, shape (M, L) C-order == Julia's L x M column-major (one sequence per column).

    Match columns = characters of the first record that are not '.' and not lowercase; a sequence is
    kept when (#gaps in match columns)/L <= max_gap_fraction (src/GaussDCA.jl:20)."""
    recs = [s for _, s in read_fasta_records(path)]
    if not recs:
        raise ValueError("no sequences in " + path)
    first = recs[0]
    inds = [i for i, c in enumerate(first) if c != "." and c == c.upper()]
    L = len(inds)
    keep = []
    for s in recs:
        if len(s) != len(first):
            raise ValueError("inputs are not aligned")
        cols = [i for i, c in enumerate(s) if c != "." and c == c.upper()]
        if cols != inds:
            raise ValueError("inconsistent inputs")
        ngaps = sum(1 for i in inds if s[i] == "-")
        if ngaps / L <= max_gap_fraction:
            keep.append(s)
    if not keep:
        raise ValueError("no sequence passed the gap filter")
    idx = np.asarray(inds)
    Z = np.empty((len(keep), L), dtype=np.int8)
    for k, s in enumerate(keep):
        Z[k] = _LETTER2NUM[np.frombuffer(s.encode("ascii"), dtype=np.uint8)[idx]]
    return Z


def remove_duplicate_sequences(Z: np.ndarray) -> np.ndarray:
    """Keep the first occurrence of every distinct sequence, order preserved (src/GaussDCA.jl:21-23)."""
    seen, keep = set(), []
    for k in range(Z.shape[0]):
        key = Z[k].tobytes()
        if key not in seen:
            seen.add(key)
            keep.append(k)
    return np.ascontiguousarray(Z[keep])


# --------------------------------------------------------------------------------------------
# theta / weights / frequencies  (DCAUtils compute_weighted_frequencies, call site src/GaussDCA.jl:28)
# --------------------------------------------------------------------------------------------
def compress_Z(Z: np.ndarray) -> np.ndarray:
    M, L = Z.shape
    nw = lib().oracle_words_per_seq(L)
    cZ = np.empty((M, nw), dtype=np.uint64)
    lib().oracle_compress_Z(_ptr(np.ascontiguousarray(Z)), L, M, _ptr(cZ))
    return cZ


def ident_sum(Z: np.ndarray, packed: bool = True) -> int:
    """sum_{k<l} #equal positions (gap==gap counts).  Exact integer."""
    M, L = Z.shape
    Z = np.ascontiguousarray(Z)
    if packed:
        cZ = compress_Z(Z)
        return int(lib().oracle_ident_sum_packed(_ptr(cZ), L, M))
    return int(lib().oracle_ident_sum_bytes(_ptr(Z), L, M))


def theta_from_ident_sum(total_ident: int, L: int, M: int) -> float:
    """theta = min(0.5, 0.38*0.32/meanfracid), meanfracid = (sum ident / L) / (M(M-1)/2)."""
    npairs = 0.5 * M * (M - 1)
    meanfracid = (total_ident / L) / npairs
    if meanfracid == 0.0:  # Julia: 0.1216/0.0 == Inf, min(0.5, Inf) == 0.5
        return 0.5
    return min(0.5, 0.38 * 0.32 / meanfracid)


def compute_theta(Z: np.ndarray, packed: bool = True) -> float:
    M, L = Z.shape
    return theta_from_ident_sum(ident_sum(Z, packed), L, M)


def meff_from_counts(counts: np.ndarray) -> float:
    """Correctly rounded sum_k 1/count[k]: exact rational sum over the count histogram."""
    vals, h = np.unique(counts, return_counts=True)
    tot = Fraction(0)
    for c, m in zip(vals.tolist(), h.tolist()):
        tot += Fraction(m, c)
    return float(tot)


def compute_weights(Z: np.ndarray, theta: float, packed: bool = True):
    """-> counts int32[M], W f64[M] (= 1/count), Meff, thresh.   theta == 0 => W == 1, Meff == M."""
    M, L = Z.shape
    if theta == 0:
        counts = np.ones(M, dtype=np.int32)
        return counts, np.ones(M), float(M), 0
    thresh = int(math.floor(theta * L))
    counts = np.empty(M, dtype=np.int32)
    Z = np.ascontiguousarray(Z)
    if packed:
        cZ = compress_Z(Z)
        lib().oracle_neighbour_counts_packed(_ptr(cZ), L, M, thresh, _ptr(counts))
    else:
        lib().oracle_neighbour_counts_bytes(_ptr(Z), L, M, thresh, _ptr(counts))
    W = 1.0 / counts.astype(np.float64)
    return counts, W, meff_from_counts(counts), thresh


def compute_freqs(Z: np.ndarray, q: int, W: np.ndarray, Meff: float):
    M, L = Z.shape
    n = (q - 1) * L
    Pi = np.empty(n)
    Pij = np.empty((n, n))
    lib().oracle_weighted_freqs(_ptr(np.ascontiguousarray(Z)), L, M, q, _ptr(np.ascontiguousarray(W)), Meff,
                                _ptr(Pi), _ptr(Pij))
    return Pi, Pij


def compute_weighted_frequencies(Z: np.ndarray, q: int, theta="auto", packed: bool = True):
    """-> Pi_true, Pij_true, Meff, W, info      (call site src/GaussDCA.jl:28)"""
    th = compute_theta(Z, packed) if theta == "auto" else float(theta)
    counts, W, Meff, thresh = compute_weights(Z, th, packed)
    Pi, Pij = compute_freqs(Z, q, W, Meff)
    return Pi, Pij, Meff, W, {"theta": th, "thresh": thresh, "counts": counts}


# --------------------------------------------------------------------------------------------
# pseudocount, covariance, inverse
# --------------------------------------------------------------------------------------------
def add_pseudocount(Pi_true, Pij_true, pc: float, q: int):
    """DCAUtils add_pseudocount, call site src/GaussDCA.jl:30.  Diagonal site blocks get
    (1-pc)*Pij_true + delta_ab*pc/q (no pc/q^2 there), the rest (1-pc)*Pij_true + pc/q^2."""
    s = q - 1
    n = Pi_true.shape[0]
    L = n // s
    pcq = pc / q
    Pi = (1 - pc) * Pi_true + pcq
    Pij = (1 - pc) * Pij_true + pcq / q
    for i in range(L):
        sl = slice(i * s, (i + 1) * s)
        blk = (1 - pc) * Pij_true[sl, sl]
        blk[np.arange(s), np.arange(s)] += pcq
        Pij[sl, sl] = blk
    return Pi, Pij


def compute_C(Pi, Pij):
    """src/GaussDCA.jl:32,76"""
    return Pij - np.outer(Pi, Pi)


class PosDefException(ArithmeticError):
    def __init__(self, info):
        super().__init__(f"matrix is not positive definite; Cholesky factorization failed (info={info})")
        self.info = info


def inv_cholesky(C):
    """mJ = inv(cholesky(C))  (src/GaussDCA.jl:34): LAPACK dpotrf('U') + dpotri + symmetrise."""
    from scipy.linalg import lapack

    c, info = lapack.dpotrf(C, lower=0, clean=0, overwrite_a=0)
    if info != 0:
        raise PosDefException(info)
    inv, info = lapack.dpotri(c, lower=0, overwrite_c=1)
    if info != 0:
        raise PosDefException(info)
    iu = np.triu_indices_from(inv, 1)
    inv.T[iu] = inv[iu]
    return inv


# --------------------------------------------------------------------------------------------
# block scores
# --------------------------------------------------------------------------------------------
def compute_FN(mJ, q: int):
    """DCAUtils compute_FN, call site src/GaussDCA.jl:39.  Gauge: means over the s x s block."""
    s = q - 1
    L = mJ.shape[0] // s
    B = mJ.reshape(L, s, L, s).transpose(0, 2, 1, 3)  # [i, j, a, b]
    K = B - B.mean(axis=3, keepdims=True) - B.mean(axis=2, keepdims=True) + B.mean(axis=(2, 3), keepdims=True)
    FN = np.sqrt((K * K).sum(axis=(2, 3)))
    FN = np.triu(FN, 1)
    return FN + FN.T


def _sqrtm_spd(A):
    w, V = np.linalg.eigh(A)
    return (V * np.sqrt(np.maximum(w, 0.0))) @ V.T


def compute_DI_gauss(mJ, C, q: int):
    """DCAUtils compute_DI_gauss, call site src/GaussDCA.jl:37.
    V = (sqrt(C_ii) mJ_ij sqrt(C_jj)) (.)^T ; DI = s/2*log(1/2) + 1/2 sum_k log(1 + sqrt(1 + 4 lambda_k(V)))."""
    s = q - 1
    L = mJ.shape[0] // s
    z = 0.5 * s * math.log(0.5)
    sq = [_sqrtm_spd(C[i * s:(i + 1) * s, i * s:(i + 1) * s]) for i in range(L)]
    DI = np.zeros((L, L))
    for i in range(L - 1):
        left = sq[i] @ mJ[i * s:(i + 1) * s, :]
        for j in range(i + 1, L):
            MM = left[:, j * s:(j + 1) * s] @ sq[j]
            lam = np.linalg.eigvalsh(MM @ MM.T)
            lam = np.maximum(lam, 0.0)
            DI[i, j] = DI[j, i] = z + 0.5 * np.sum(np.log(1.0 + np.sqrt(1.0 + 4.0 * lam)))
    return DI


def correct_APC(S):
    """src/GaussDCA.jl:78-86 (diagonal zeros take part in the sums)."""
    N = S.shape[0]
    Si = S.sum(axis=0, keepdims=True)
    Sj = S.sum(axis=1, keepdims=True)
    Sa = S.sum() * (1 - 1 / N)
    return S - (Sj @ Si) / Sa


def compute_ranking(S, min_separation: int = 5):
    """src/GaussDCA.jl:88-99: enumerate i<j, j-i>=min_separation, value S[j,i]; stable sort descending.
    -> list of (i, j, score), 1-based."""
    N = S.shape[0]
    ii, jj, vv = [], [], []
    for i in range(N - min_separation):
        j = np.arange(i + min_separation, N)
        ii.append(np.full(j.shape, i))
        jj.append(j)
        vv.append(S[j, i])
    if not ii:
        return []
    ii, jj, vv = np.concatenate(ii), np.concatenate(jj), np.concatenate(vv)
    # Julia sorts with isless (a total order: -0.0 < 0.0), stable, rev=true.  Map doubles to integers
    # that sort the same way, complement for descending, and let a stable argsort keep enumeration order on ties.
    u = np.ascontiguousarray(vv, dtype=np.float64).view(np.uint64).copy()
    u[np.isnan(vv)] = np.uint64(0x7FF8000000000000)  # isless: any NaN (either sign bit) is above +Inf, NaNs tie
    asc = np.where((u >> np.uint64(63)).astype(bool), ~u, u | np.uint64(1 << 63))
    order = np.argsort(~asc, kind="stable")
    return [(int(ii[o]) + 1, int(jj[o]) + 1, float(vv[o])) for o in order]


def check_arguments(filename, pseudocount, theta, max_gap_fraction, score, min_separation):
    """src/GaussDCA.jl:49-65"""
    if not (0 <= pseudocount <= 1):
        raise ValueError(f"invalid pseudocount value: {pseudocount} (must be between 0 and 1)")
    if not (theta == "auto" or (isinstance(theta, (int, float)) and 0 <= theta <= 1)):
        raise ValueError(f"invalid θ value: {theta} (must be either :auto, or a number between 0 and 1)")
    if not (0 <= max_gap_fraction <= 1):
        raise ValueError(f"invalid max_gap_fraction value: {max_gap_fraction} (must be between 0 and 1)")
    if score not in ("DI", "frob"):
        raise ValueError(f"invalid score value: {score} (must be either :DI or :frob)")
    if not (min_separation >= 1):
        raise ValueError(f"invalid min_separation value: {min_separation} (must be >= 1)")
    if not os.path.isfile(filename):
        raise ValueError(f"cannot open file {filename}")
    return True


def gdca_from_Z(Z, pseudocount=0.8, theta="auto", score="frob", min_separation=5, packed=True, stages=None):
    """src/GaussDCA.jl:24-46 from the encoded alignment onward."""
    q = int(Z.max())
    if q >= 32:
        raise RuntimeError(f"parameter q={q} is too big (max 31 is allowed)")
    Pi_true, Pij_true, Meff, W, info = compute_weighted_frequencies(Z, q, theta, packed)
    Pi, Pij = add_pseudocount(Pi_true, Pij_true, float(pseudocount), q)
    C = compute_C(Pi, Pij)
    mJ = inv_cholesky(C)
    S = compute_DI_gauss(mJ, C, q) if score == "DI" else compute_FN(mJ, q)
    Sc = correct_APC(S)
    R = compute_ranking(Sc, min_separation)
    if stages is not None:
        stages.update(q=q, Meff=Meff, W=W, C=C, mJ=mJ, S_raw=S, S=Sc, **info)
    return R


def gDCA(filename, pseudocount=0.8, theta="auto", max_gap_fraction=0.9, score="frob", min_separation=5,
         remove_dups=False, packed=True, stages=None):
    """src/GaussDCA.jl:8-47"""
    check_arguments(filename, pseudocount, theta, max_gap_fraction, score, min_separation)
    Z = read_fasta_alignment(filename, max_gap_fraction)
    if remove_dups:
        Z = remove_duplicate_sequences(Z)
    return gdca_from_Z(Z, pseudocount, theta, score, min_separation, packed, stages)


def format_rank(R) -> str:
    """printrank, src/GaussDCA.jl:67-70: '%i %i %e\\n'"""
    def e(x):  # Julia's @printf spells the non-finite values NaN / Inf / -Inf
        return "%e" % x if math.isfinite(x) else ("NaN" if math.isnan(x) else ("Inf" if x > 0 else "-Inf"))
    return "".join("%i %i %s\n" % (i, j, e(x)) for i, j, x in R)


def synth_alignment(L: int, M: int, seed: int = 20140321) -> np.ndarray:
    """Clustered synthetic alignment of SURVEY 8(d); -> int8 (M, L)."""
    Z = np.empty((M, L), dtype=np.int8)
    lib().oracle_synth_alignment(_ptr(Z), L, M, seed)
    return Z
