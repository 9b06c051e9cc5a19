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


class _HostBlock:
    """Owner of a block returned by the C front-end (gdca_free_host on collection)."""

    def __init__(self, lib, address):
        self._lib, self._address = lib, address

    def __del__(self):
        if self._address:
            self._lib.gdca_free_host(ctypes.c_void_p(self._address))
            self._address = 0


def read_fasta_alignment(filename, max_gap_fraction: float) -> np.ndarray:
    """-> Z int8, shape (M, L), C-contiguous: Z[k] is sequence k.  Memory-identical to the reference's
    L x M column-major Matrix{Int8} (src/GaussDCA.jl:20,24)."""
    lib = _lib.load()
    zp, L, M = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int64()
    st = lib.gdca_read_fasta_alignment(os.fsencode(filename), float(max_gap_fraction), ctypes.byref(zp),
                                       ctypes.byref(L), ctypes.byref(M))
    if st != _lib.GDCA_OK:
        raise ValueError(_host_error(lib))
    # zero-copy: the array views the block the C reader allocated; the block is released when the last view is gone
    buf = (ctypes.c_int8 * (L.value * M.value)).from_address(zp.value)
    buf._owner = _HostBlock(lib, zp.value)            # numpy keeps `buf` alive as the array's base object
    return np.frombuffer(buf, dtype=np.int8).reshape(M.value, L.value)


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
