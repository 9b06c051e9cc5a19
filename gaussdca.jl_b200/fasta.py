"""Host I/O that stays on the CPU (north star: "FASTA parsing, gap filtering and deduplication stay on
the host as I/O"): reference src/GaussDCA.jl:20-23, i.e. DCAUtils read_fasta_alignment /
remove_duplicate_sequences (un-vendored).  numpy-vectorised; independent of oracle/."""
from __future__ import annotations

import gzip

import numpy as np

# A..Y -> 1..20 for the standard amino acids; B, J, O, U, X, Z, '-', everything else -> 21
_CODE = np.full(256, 21, dtype=np.int8)
for _k, _ch in enumerate("ACDEFGHIKLMNPQRSTVWY"):
    _CODE[ord(_ch)] = _k + 1
_IS_MATCH = np.ones(256, dtype=bool)  # match column: not '.', not a lowercase letter
_IS_MATCH[ord(".")] = False
_IS_MATCH[ord("a"):ord("z") + 1] = False


def _records(path):
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rb") as fh:
        data = fh.read()
    out = []
    for rec in data.split(b">")[1:]:
        nl = rec.find(b"\n")
        body = rec[nl + 1:] if nl >= 0 else b""
        out.append(body.translate(None, b" \t\r\n"))
    return out


def read_fasta_alignment(filename, max_gap_fraction: float) -> np.ndarray:
    """-> Z int8, shape (M, L), C-contiguous: Z[k] is sequence k.  Memory-identical to the reference's
    L x M column-major Matrix{Int8} (src/GaussDCA.jl:20,24)."""
    seqs = _records(filename)
    if not seqs:
        raise ValueError(f"no sequences found in {filename}")
    first = np.frombuffer(seqs[0], dtype=np.uint8)
    cols = np.flatnonzero(_IS_MATCH[first])
    L = cols.size
    if L == 0:
        raise ValueError("alignment has no match columns")
    rows = []
    for s in seqs:
        a = np.frombuffer(s, dtype=np.uint8)
        if a.size != first.size:
            raise ValueError("inputs are not aligned")
        if not np.array_equal(np.flatnonzero(_IS_MATCH[a]), cols):
            raise ValueError("inconsistent inputs")
        m = a[cols]
        if np.count_nonzero(m == ord("-")) / L <= max_gap_fraction:
            rows.append(_CODE[m])
    if not rows:
        raise ValueError(f"Out of {len(seqs)} sequences, none passed the filter (max_gap_fraction={max_gap_fraction})")
    return np.ascontiguousarray(np.stack(rows))


def remove_duplicate_sequences(Z: np.ndarray):
    """-> (Znew, kept indices): first occurrence of each distinct sequence, order kept (src/GaussDCA.jl:21-23)."""
    Z = np.ascontiguousarray(Z)
    v = Z.view(np.dtype((np.void, Z.shape[1])))[:, 0]
    _, first = np.unique(v, return_index=True)
    keep = np.sort(first)
    return np.ascontiguousarray(Z[keep]), keep
