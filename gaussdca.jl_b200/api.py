"""Host-side mirror of the reference's public surface (src/GaussDCA.jl:3 exports gDCA, printrank).

Same names, keyword meaning, defaults and error behaviour as the Julia module; the body between the
encoded alignment (src/GaussDCA.jl:24) and the ranking (:44) is one call into libgdca_b200.so.
Julia is not installed in this image, so this Python layer is the executable stand-in for the
Julia wrapper in julia/GaussDCA.jl (same ccall sequence).

Array convention: Z is int8 of shape (M, L), C-contiguous, Z[k] = sequence k -- byte-identical to
Julia's L x M column-major Matrix{Int8}.  Symmetric matrices are returned full.
"""
from __future__ import annotations

import ctypes
import os
import sys

import numpy as np

from . import _lib
from ._lib import RANK_DTYPE, SCORE_CODES, PosDefException, default_context, ptr
from .fasta import read_fasta_alignment, remove_duplicate_sequences

__all__ = ["gDCA", "printrank", "check_arguments", "gdca_from_alignment", "compute_theta", "compute_weights",
           "compute_covariance", "compute_weighted_frequencies", "add_pseudocount", "compute_C", "inverse", "compute_FN", "compute_DI_gauss", "correct_APC", "compute_ranking",
           "read_fasta_alignment", "remove_duplicate_sequences", "PosDefException"]


def _theta_code(theta) -> float:
    return -1.0 if (isinstance(theta, str) and theta == "auto") else float(theta)


def check_arguments(filename, pseudocount, theta, max_gap_fraction, score, min_separation):
    """src/GaussDCA.jl:49-65 -- same checks, same messages; ArgumentError -> ValueError."""
    def aerror(s):
        raise ValueError(s)
    if not (0 <= pseudocount <= 1):
        aerror(f"invalid pseudocount value: {pseudocount} (must be between 0 and 1)")
    is_auto = isinstance(theta, str) and theta == "auto"
    is_real = isinstance(theta, (int, float, np.integer, np.floating)) and not isinstance(theta, bool)
    if not (is_auto or (is_real and 0 <= theta <= 1)):
        aerror(f"invalid θ value: {theta} (must be either :auto, or a number between 0 and 1)")
    if not (0 <= max_gap_fraction <= 1):
        aerror(f"invalid max_gap_fraction value: {max_gap_fraction} (must be between 0 and 1)")
    if score not in ("DI", "frob"):
        aerror(f"invalid score value: {score} (must be either :DI or :frob)")
    if not (min_separation >= 1):
        aerror(f"invalid min_separation value: {min_separation} (must be >= 1)")
    if not os.path.isfile(filename):
        aerror(f"cannot open file {filename}")
    return True


def _as_Z(Z) -> np.ndarray:
    Z = np.ascontiguousarray(Z, dtype=np.int8)
    if Z.ndim != 2:
        raise ValueError("Z must be a 2-D int8 array of shape (M, L)")
    return Z


def gdca_from_alignment(Z, pseudocount=0.8, theta="auto", score="frob", min_separation=5, *, ctx=None,
                        return_stats=False, as_array=False):
    """src/GaussDCA.jl:24-46 on an already encoded alignment: one gdca_run() call."""
    ctx = ctx or default_context()
    Z = _as_Z(Z)
    M, L = Z.shape
    n_out = int(ctx.lib.gdca_ranking_length(L, int(min_separation)))
    R = np.empty(n_out, dtype=RANK_DTYPE)
    st = _lib.Stats()
    code = ctx.lib.gdca_run(ctx.h, ptr(Z), L, M, _theta_code(theta), float(pseudocount), SCORE_CODES[score],
                            int(min_separation), ptr(R), n_out, ctypes.byref(st))
    ctx.check(code)
    out = R if as_array else [(int(i), int(j), float(x)) for i, j, x in R.tolist()]
    return (out, st.asdict()) if return_stats else out


def gDCA(filename, pseudocount=0.8, theta="auto", max_gap_fraction=0.9, score="frob", min_separation=5,
         remove_dups=False, *, θ=None, ctx=None, return_stats=False, as_array=False):
    """Drop-in for GaussDCA.gDCA (src/GaussDCA.jl:8-47).  Returns [(i, j, score), ...] sorted by
    descending score, i < j 1-based, j - i >= min_separation."""
    if θ is not None:
        theta = θ
    if isinstance(score, str) and score.startswith(":"):
        score = score[1:]
    if isinstance(theta, str) and theta == ":auto":
        theta = "auto"
    check_arguments(filename, pseudocount, theta, max_gap_fraction, score, min_separation)
    Z = read_fasta_alignment(filename, max_gap_fraction)   # host I/O, src/GaussDCA.jl:20
    if remove_dups:
        Z, _ = remove_duplicate_sequences(Z)               # host I/O, src/GaussDCA.jl:21-23
    return gdca_from_alignment(Z, pseudocount, theta, score, min_separation, ctx=ctx, return_stats=return_stats,
                               as_array=as_array)


def printrank(*args):
    """printrank(io, R) / printrank(R) / printrank(outfile, R): '%i %i %e' per row (src/GaussDCA.jl:67-74).
    The 1-argument form writes to stdout (the reference names the long-gone STDOUT there)."""
    if len(args) == 1:
        io, R = sys.stdout, args[0]
    elif len(args) == 2:
        io, R = args
    else:
        raise TypeError("printrank([io|outfile,] R)")
    arr = R if isinstance(R, np.ndarray) else np.array(R, dtype=RANK_DTYPE).reshape(-1)
    arr = np.ascontiguousarray(arr, dtype=RANK_DTYPE)
    lib = _lib.load()
    if isinstance(io, (str, os.PathLike)):
        if lib.gdca_write_rank(os.fsencode(io), ptr(arr), arr.size) != _lib.GDCA_OK:
            raise OSError(lib.gdca_host_last_error().decode())
        return
    buf = ctypes.create_string_buffer(64 * arr.size + 1)
    used = ctypes.c_int64()
    if lib.gdca_format_rank(ptr(arr), arr.size, buf, len(buf), ctypes.byref(used)) != _lib.GDCA_OK:
        raise RuntimeError(lib.gdca_host_last_error().decode())
    io.write(buf.raw[:used.value].decode("ascii"))


# ------------------------------------------------------------------ staged, DCAUtils-shaped pieces
def compute_weights(Z, theta="auto", *, ctx=None, full=False):
    """DCAUtils compute_weights (+ compute_theta when theta == 'auto'); call site src/GaussDCA.jl:28.
    -> (W, Meff);  full=True -> dict with counts, W, Meff, theta, thresh, ident_sum."""
    ctx = ctx or default_context()
    Z = _as_Z(Z)
    M, L = Z.shape
    counts = np.empty(M, dtype=np.int32)
    W = np.empty(M, dtype=np.float64)
    meff, th = ctypes.c_double(), ctypes.c_double()
    thresh, ident = ctypes.c_int64(), ctypes.c_uint64()
    ctx.check(ctx.lib.gdca_compute_weights(ctx.h, ptr(Z), L, M, _theta_code(theta), ptr(counts), ptr(W),
                                           ctypes.byref(meff), ctypes.byref(th), ctypes.byref(thresh),
                                           ctypes.byref(ident)))
    if full:
        return dict(counts=counts, W=W, Meff=meff.value, theta=th.value, thresh=thresh.value,
                    ident_sum=ident.value, passes=ctx.stats()["theta_passes"])
    return W, meff.value


def compute_theta(Z, *, ctx=None) -> float:
    """DCAUtils compute_theta: min(0.5, 0.38*0.32/meanfracid)."""
    return compute_weights(Z, "auto", ctx=ctx, full=True)["theta"]


def compute_covariance(Z, W, Meff, pseudocount, *, ctx=None):
    """compute_freqs + add_pseudocount + compute_C (src/GaussDCA.jl:28-32) -> (C, Pi, q)."""
    ctx = ctx or default_context()
    Z = _as_Z(Z)
    M, L = Z.shape
    q = int(Z.max())
    if q >= 32:
        raise RuntimeError(f"parameter q={q} is too big (max 31 is allowed)")
    n = (q - 1) * L
    C = np.empty((n, n), dtype=np.float64)
    Pi = np.empty(n, dtype=np.float64)
    W = np.ascontiguousarray(W, dtype=np.float64)
    qo = ctypes.c_int32()
    ctx.check(ctx.lib.gdca_compute_covariance(ctx.h, ptr(Z), L, M, ptr(W), float(Meff), float(pseudocount), ptr(C),
                                              ptr(Pi), ctypes.byref(qo)))
    return C, Pi, qo.value


def compute_weighted_frequencies(Z, q=None, theta="auto", *, ctx=None):
    """DCAUtils compute_weighted_frequencies(Z, q, θ) -> (Pi_true, Pij_true, Meff, W); call site src/GaussDCA.jl:28.
    q must equal max(Z) (the reference passes exactly that, src/GaussDCA.jl:25) or be omitted."""
    ctx = ctx or default_context()
    Z = _as_Z(Z)
    M, L = Z.shape
    qz = int(Z.max())
    if q is not None and int(q) != qz:
        raise ValueError(f"q={q} does not match max(Z)={qz}")
    if qz >= 32:
        raise RuntimeError(f"parameter q={qz} is too big (max 31 is allowed)")
    n = (qz - 1) * L
    Pi = np.empty(n, dtype=np.float64)
    Pij = np.empty((n, n), dtype=np.float64)
    W = np.empty(M, dtype=np.float64)
    meff, th, qq = ctypes.c_double(), ctypes.c_double(), ctypes.c_int32()
    ctx.check(ctx.lib.gdca_compute_weighted_frequencies(ctx.h, ptr(Z), L, M, _theta_code(theta), ptr(Pi), ptr(Pij),
                                                        ctypes.byref(meff), ptr(W), ctypes.byref(th), ctypes.byref(qq)))
    return Pi, Pij, meff.value, W


def add_pseudocount(Pi_true, Pij_true, pseudocount, q, *, ctx=None):
    """DCAUtils add_pseudocount(Pi_true, Pij_true, pc, q) -> (Pi, Pij); call site src/GaussDCA.jl:30."""
    ctx = ctx or default_context()
    Pi_true = np.ascontiguousarray(Pi_true, dtype=np.float64)
    Pij_true = np.ascontiguousarray(Pij_true, dtype=np.float64)
    n = Pi_true.shape[0]
    if Pij_true.shape != (n, n):
        raise ValueError("Pij_true must be n x n with n = len(Pi_true)")
    Pi = np.empty(n, dtype=np.float64)
    Pij = np.empty((n, n), dtype=np.float64)
    ctx.check(ctx.lib.gdca_add_pseudocount(ctx.h, ptr(Pi_true), ptr(Pij_true), n, int(q), float(pseudocount), ptr(Pi), ptr(Pij)))
    return Pi, Pij


def compute_C(Pi, Pij, *, ctx=None):
    """compute_C(Pi, Pij) = Pij - Pi * Pi' (src/GaussDCA.jl:32,76)."""
    ctx = ctx or default_context()
    Pi = np.ascontiguousarray(Pi, dtype=np.float64)
    Pij = np.ascontiguousarray(Pij, dtype=np.float64)
    n = Pi.shape[0]
    if Pij.shape != (n, n):
        raise ValueError("Pij must be n x n with n = len(Pi)")
    C = np.empty((n, n), dtype=np.float64)
    ctx.check(ctx.lib.gdca_compute_C(ctx.h, ptr(Pi), ptr(Pij), n, ptr(C)))
    return C


def inverse(C, *, ctx=None):
    """mJ = inv(cholesky(C)) (src/GaussDCA.jl:34); raises PosDefException(info) like the reference."""
    ctx = ctx or default_context()
    C = np.ascontiguousarray(C, dtype=np.float64)
    n = C.shape[0]
    mJ = np.empty_like(C)
    info = ctypes.c_int32()
    ctx.check(ctx.lib.gdca_inverse(ctx.h, ptr(C), n, ptr(mJ), ctypes.byref(info)))
    return mJ


def _score(mJ, C, q, which, ctx):
    ctx = ctx or default_context()
    mJ = np.ascontiguousarray(mJ, dtype=np.float64)
    n = mJ.shape[0]
    L = n // (q - 1)
    S = np.empty((L, L), dtype=np.float64)
    Cc = None if C is None else np.ascontiguousarray(C, dtype=np.float64)
    ctx.check(ctx.lib.gdca_score(ctx.h, ptr(mJ), ptr(Cc), n, int(q), SCORE_CODES[which], ptr(S)))
    return S


def compute_FN(mJ, q, *, ctx=None):
    """DCAUtils compute_FN (call site src/GaussDCA.jl:39)."""
    return _score(mJ, None, q, "frob", ctx)


def compute_DI_gauss(mJ, C, q, *, ctx=None):
    """DCAUtils compute_DI_gauss (call site src/GaussDCA.jl:37)."""
    return _score(mJ, C, q, "DI", ctx)


def correct_APC(S, *, ctx=None):
    """src/GaussDCA.jl:78-86"""
    ctx = ctx or default_context()
    S = np.ascontiguousarray(S, dtype=np.float64)
    out = np.empty_like(S)
    ctx.check(ctx.lib.gdca_apc(ctx.h, ptr(S), S.shape[0], ptr(out)))
    return out


def compute_ranking(S, min_separation=5, *, ctx=None, as_array=False):
    """src/GaussDCA.jl:88-99"""
    ctx = ctx or default_context()
    S = np.ascontiguousarray(S, dtype=np.float64)
    L = S.shape[0]
    n_out = int(ctx.lib.gdca_ranking_length(L, int(min_separation)))
    R = np.empty(n_out, dtype=RANK_DTYPE)
    ctx.check(ctx.lib.gdca_ranking(ctx.h, ptr(S), L, int(min_separation), ptr(R), n_out))
    return R if as_array else [(int(i), int(j), float(x)) for i, j, x in R.tolist()]
