"""Host I/O that stays on the CPU (north star: "FASTA parsing, gap filtering and deduplication stay on
the host as I/O"): reference src/GaussDCA.jl:20-23, i.e. DCAUtils read_fasta_alignment /
remove_duplicate_sequences (un-vendored).  Thin ctypes wrappers over the C++ front-end of
libgdca_b200.so (csrc/host_io.cpp: one pass, zlib, OpenMP) -- no GPU needed; independent of oracle/."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib


def _host_error(lib) -> str:
    return lib.gdca_host_last_error().decode()


def read_fasta_alignment(filename, max_gap_fraction: float) -> np.ndarray:
    """-> Z int8, shape (M, L), C-contiguous: Z[k] is sequence k.  Memory-identical to the reference's
    L x M column-major Matrix{Int8} (src/GaussDCA.jl:20,24)."""
    lib = _lib.load()
    zp, L, M = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int64()
    st = lib.gdca_read_fasta_alignment(os.fsencode(filename), float(max_gap_fraction), ctypes.byref(zp),
                                       ctypes.byref(L), ctypes.byref(M))
    if st != _lib.GDCA_OK:
        raise ValueError(_host_error(lib))
    try:
        buf = (ctypes.c_int8 * (L.value * M.value)).from_address(zp.value)
        return np.frombuffer(buf, dtype=np.int8).reshape(M.value, L.value).copy()
    finally:
        lib.gdca_free_host(zp)


def remove_duplicate_sequences(Z: np.ndarray):
    """-> (Znew, kept indices): first occurrence of each distinct sequence, order kept (src/GaussDCA.jl:21-23)."""
    lib = _lib.load()
    Z = np.ascontiguousarray(Z, dtype=np.int8)
    M, L = Z.shape
    out = np.empty_like(Z)
    kept = np.empty(M, dtype=np.int64)
    m = ctypes.c_int64()
    st = lib.gdca_remove_duplicate_sequences(_lib.ptr(Z), L, M, _lib.ptr(out), ctypes.byref(m), _lib.ptr(kept))
    if st != _lib.GDCA_OK:
        raise ValueError(_host_error(lib))
    return np.ascontiguousarray(out[:m.value]), kept[:m.value].copy()
