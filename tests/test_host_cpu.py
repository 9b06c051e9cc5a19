"""CPU: host-side logic of the product and the C-ABI surface (no GPU, no compute calls)."""
import ctypes
import io
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_path


def header_functions():
    src = open(os.path.join(ROOT, "include", "gdca_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gdca_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    """The in-tree .so loads and exports everything include/gdca_b200.h declares; the ctypes table covers it."""
    from gaussdca_jl_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = pkg.load()
    names = header_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} declared in the header but not bound in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.gdca_abi_version() == 2
    assert lib.gdca_ranking_length(53, 5) == 48 * 49 // 2 == 1176           # src/GaussDCA.jl:90
    assert lib.gdca_ranking_length(53, 4) == 1225 and lib.gdca_ranking_length(400, 5) == 78210
    assert lib.gdca_ranking_length(5, 5) == 0 and lib.gdca_ranking_length(3, 5) == 0
    assert lib.gdca_status_string(3) == b"matrix is not positive definite"


def test_struct_layouts_match_the_header(pkg, tmp_path):
    from gaussdca_jl_b200 import _lib
    c = tmp_path / "sz.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gdca_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu",'
                 'sizeof(gdca_rank_t),offsetof(gdca_rank_t,j),offsetof(gdca_rank_t,score),sizeof(gdca_stats_t),'
                 'offsetof(gdca_stats_t,theta),offsetof(gdca_stats_t,ms_h2d));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    S = _lib.Stats
    assert got == [24, 8, 16, ctypes.sizeof(S), S.theta.offset, S.ms_h2d.offset]
    assert _lib.RANK_DTYPE.itemsize == 24 and _lib.RANK_DTYPE.fields["j"][1] == 8 and _lib.RANK_DTYPE.fields["score"][1] == 16


def test_theta_host_arithmetic_matches_oracle(pkg, orc):
    """gdca_theta_from_ham_sum is host code in the library: same IEEE operations as the oracle."""
    from gaussdca_jl_b200.dist import theta_from_ham, theta_from_ident
    lib = pkg.load()
    rng = np.random.default_rng(0)
    for L, M in [(53, 106), (400, 94), (500, 200000), (1500, 1000000), (1, 2)]:
        npairs = M * (M - 1) // 2
        for frac in (0.0, 0.05, 0.3123, 0.9, 1.0):
            ident = int(npairs * L * frac) + int(rng.integers(0, 3)) * (frac not in (0.0, 1.0))
            ham = npairs * L - ident
            th, thr, ids = ctypes.c_double(), ctypes.c_int64(), ctypes.c_uint64()
            assert lib.gdca_theta_from_ham_sum(L, M, ham, ctypes.byref(th), ctypes.byref(thr), ctypes.byref(ids)) == 0
            want = orc.theta_from_ident_sum(ident, L, M)
            assert th.value == want and ids.value == ident and thr.value == int(np.floor(want * L))
            assert theta_from_ham(L, M, ham) == (want, thr.value, ident)
            assert theta_from_ident(L, M, ident) == (want, thr.value)
            th2, thr2 = ctypes.c_double(), ctypes.c_int64()
            assert lib.gdca_theta_from_ident_sum(L, M, ident, ctypes.byref(th2), ctypes.byref(thr2)) == 0
            assert (th2.value, thr2.value) == (want, thr.value)


def test_no_cpu_fallback(pkg):
    """Without a GPU the product refuses to run (it must never route through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.GdcaError, match="no CPU fallback"):
        pkg.Context(0)
    with pytest.raises(pkg.GdcaError):
        pkg.gDCA(golden_path("small.fasta.gz"))
    src = "".join(open(os.path.join(ROOT, "gaussdca.jl_b200", f)).read()
                  for f in ("__init__.py", "api.py", "_lib.py", "fasta.py", "dist.py"))
    assert "oracle" not in src.replace("oracle/", "").replace("the oracle", "").replace("oracle's", "").replace("oracle theta", "")


def test_check_arguments_messages(pkg):
    """src/GaussDCA.jl:49-65: same conditions, same messages."""
    fa = golden_path("small.fasta.gz")
    ok = dict(filename=fa, pseudocount=0.8, theta="auto", max_gap_fraction=0.9, score="frob", min_separation=5)
    assert pkg.check_arguments(**ok) is True
    cases = [
        (dict(pseudocount=-0.1), "invalid pseudocount value: -0.1 (must be between 0 and 1)"),
        (dict(theta=1.5), "invalid θ value: 1.5 (must be either :auto, or a number between 0 and 1)"),
        (dict(theta="automatic"), "invalid θ value: automatic"),
        (dict(max_gap_fraction=2), "invalid max_gap_fraction value: 2 (must be between 0 and 1)"),
        (dict(score="mi"), "invalid score value: mi (must be either :DI or :frob)"),
        (dict(min_separation=0), "invalid min_separation value: 0 (must be >= 1)"),
        (dict(filename="/nonexistent.fasta"), "cannot open file /nonexistent.fasta"),
    ]
    for kw, msg in cases:
        with pytest.raises(ValueError) as e:
            pkg.check_arguments(**{**ok, **kw})
        assert msg in str(e.value)
    for th in (0, 0.0, 1, 0.37):
        assert pkg.check_arguments(**{**ok, "theta": th})


def test_fasta_reader_and_dedup_match_oracle(pkg, orc, tmp_path):
    for fa, mg in [("small.fasta.gz", 0.9), ("small.fasta.gz", 0.8), ("large.fasta.gz", 0.9), ("large.fasta.gz", 0.5)]:
        Z, Zo = pkg.read_fasta_alignment(golden_path(fa), mg), orc.read_fasta_alignment(golden_path(fa), mg)
        assert Z.dtype == np.int8 and Z.flags.c_contiguous and np.array_equal(Z, Zo)
        Zd, keep = pkg.remove_duplicate_sequences(Z)
        assert np.array_equal(Zd, orc.remove_duplicate_sequences(Zo)) and np.array_equal(Zd, Z[keep])
    assert pkg.read_fasta_alignment(golden_path("small.fasta.gz"), 0.9).shape == (106, 53)
    Zl = pkg.read_fasta_alignment(golden_path("large.fasta.gz"), 0.9)
    assert Zl.shape == (97, 400) and pkg.remove_duplicate_sequences(Zl)[0].shape == (94, 400)
    # insert columns ('.' and lowercase) are skipped; odd letters map to 21; plain-text files work
    p = tmp_path / "t.fasta"
    p.write_text(">a\nAC.dE-\n>b\nXZ.fBY\n>c\n--.g--\n")
    Z = pkg.read_fasta_alignment(str(p), 0.9)
    assert Z.tolist() == [[1, 2, 4, 21], [21, 21, 21, 20]]            # third sequence is all gaps -> filtered
    assert np.array_equal(Z, orc.read_fasta_alignment(str(p), 0.9))
    assert pkg.read_fasta_alignment(str(p), 1.0).shape == (3, 4)
    p.write_text(">a\nACD\n>b\nAC\n")
    with pytest.raises(ValueError, match="not aligned"):
        pkg.read_fasta_alignment(str(p), 0.9)
    p.write_text(">a\nAcD\n>b\nACd\n")
    with pytest.raises(ValueError, match="inconsistent inputs"):
        pkg.read_fasta_alignment(str(p), 0.9)
    with pytest.raises(ValueError, match="cannot open file"):
        pkg.read_fasta_alignment(str(tmp_path / "missing.fasta"), 0.9)
    p.write_text(">a\n---\n>b\n---\n")
    with pytest.raises(ValueError, match="none passed the filter"):
        pkg.read_fasta_alignment(str(p), 0.5)
    # header lines may contain '>', lines may be wrapped, CRLF line ends, blank lines, text before the first record
    p.write_bytes(b"; comment\r\n>s1 desc >still header\r\nAC\r\nDE\r\n\r\n>s2\r\nWY-K\r\n")
    Z = pkg.read_fasta_alignment(str(p), 0.9)
    assert Z.tolist() == [[1, 2, 3, 4], [19, 20, 21, 9]] and np.array_equal(Z, orc.read_fasta_alignment(str(p), 0.9))


def test_fasta_reader_large_synthetic_and_gzip(pkg, orc, tmp_path):
    """Round trip of a 3000 x 300 synthetic alignment through FASTA text (plain, wrapped at 60, and gzipped)."""
    import gzip
    L, M = 300, 3000
    Z = orc.synth_alignment(L, M, seed=4)
    letters = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY-", dtype=np.uint8)
    rows = [letters[Z[k] - 1].tobytes() for k in range(M)]
    plain = tmp_path / "a.fasta"
    with open(plain, "wb") as f:
        for k, r in enumerate(rows):
            f.write(b">seq%d\n" % k)
            for o in range(0, L, 60):
                f.write(r[o:o + 60] + b"\n")
    gz = tmp_path / "a.fasta.gz"
    with open(plain, "rb") as f, gzip.open(gz, "wb") as o:
        o.write(f.read())
    for path in (plain, gz):
        Zr = pkg.read_fasta_alignment(str(path), 1.0)
        assert np.array_equal(Zr, Z)
    # duplicates: append copies in scrambled positions, keep-first semantics and the kept indices
    Zd = np.concatenate([Z[:50], Z[10:30], Z[:5], Z[50:60]])
    out, kept = pkg.remove_duplicate_sequences(Zd)
    assert np.array_equal(out, orc.remove_duplicate_sequences(Zd)) and np.array_equal(out, Zd[kept])
    assert out.shape[0] == 60 and kept.tolist() == list(range(50)) + list(range(75, 85))


def test_printrank_format(pkg, orc, tmp_path):
    R = [(11, 35, 3.6494745366789094), (9, 37, 1.676179e+00), (1, 6, -2.5e-05)]
    buf = io.StringIO()
    pkg.printrank(buf, R)
    assert buf.getvalue() == "11 35 3.649475e+00\n9 37 1.676179e+00\n1 6 -2.500000e-05\n" == orc.format_rank(R)
    out = tmp_path / "r.txt"
    pkg.printrank(str(out), R)
    assert out.read_text() == buf.getvalue()
    big = [(i, i + 7, (-1) ** i * 1.2345678e-3 * 10.0 ** (i % 40 - 20)) for i in range(1, 3000)]
    b2 = io.StringIO()
    pkg.printrank(b2, big)
    assert b2.getvalue() == orc.format_rank(big)
    from gaussdca_jl_b200._lib import RANK_DTYPE
    arr = np.array(R, dtype=RANK_DTYPE)
    buf2 = io.StringIO()
    pkg.printrank(buf2, arr)
    assert buf2.getvalue() == buf.getvalue()


def test_julia_wrapper_is_consistent_with_header():
    """Julia is absent here, so the wrapper is checked statically: every ccall names an exported symbol and
    the public surface of src/GaussDCA.jl:3,8-16,67-74 is present."""
    jl = open(os.path.join(ROOT, "gaussdca.jl_b200", "julia", "GaussDCA.jl")).read()
    called = set(re.findall(r"ccall\(\(:(gdca_[A-Za-z0-9_]+)", jl))
    assert called and called <= set(header_functions())
    assert {"gdca_create", "gdca_run", "gdca_last_error", "gdca_ranking_length"} <= called
    assert "export gDCA, printrank" in jl
    for kw in ("pseudocount::Real = 0.8", "θ = :auto", "max_gap_fraction::Real = 0.9", "score::Symbol = :frob",
               "min_separation::Integer = 5", "remove_dups::Bool = false"):
        assert kw in jl
    assert "Vector{Tuple{Int,Int,Float64}}" in jl and "PosDefException" in jl


def test_rank_writer_is_byte_identical_across_chunk_boundaries(pkg, tmp_path):
    """gdca_write_rank / gdca_format_rank format rows in parallel chunks of 4096: the bytes must be those of the serial
    "%i %i %e\\n" loop of printrank (src/GaussDCA.jl:67-74) for any length, including inf / nan / signed zero."""
    import ctypes
    from gaussdca_jl_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(1)
    n = 3 * 4096 + 17
    R = np.zeros(n, dtype=_lib.RANK_DTYPE)
    R["i"] = rng.integers(1, 70000, n)
    R["j"] = rng.integers(1, 70000, n)
    R["score"] = rng.standard_normal(n) * 10.0 ** rng.integers(-300, 300, n)
    R["score"][:5] = [0.0, -0.0, np.inf, -np.inf, np.nan]
    path = str(tmp_path / "rank.txt")
    for m in (0, 1, 4095, 4096, 4097, 8192, n):
        def jl(x):   # Julia's @printf("%e") spells the non-finite values NaN / Inf / -Inf
            return "%e" % x if np.isfinite(x) else ("NaN" if np.isnan(x) else ("Inf" if x > 0 else "-Inf"))
        want = "".join("%d %d %s\n" % (i, j, jl(x)) for i, j, x in zip(R["i"][:m].tolist(), R["j"][:m].tolist(), R["score"][:m].tolist()))
        assert lib.gdca_write_rank(path.encode(), _lib.ptr(R), m) == 0
        assert open(path).read() == want, m
        used = ctypes.c_int64()
        assert lib.gdca_format_rank(_lib.ptr(R), m, None, 0, ctypes.byref(used)) == 0 and used.value == len(want)
        buf = ctypes.create_string_buffer(max(1, used.value))
        assert lib.gdca_format_rank(_lib.ptr(R), m, buf, used.value, ctypes.byref(used)) == 0
        assert buf.raw[: used.value].decode() == want
        if m > 1:
            assert lib.gdca_format_rank(_lib.ptr(R), m, buf, used.value - 1, ctypes.byref(used)) != 0   # buffer too small


def test_fasta_reader_fuzz_against_oracle(pkg, orc, tmp_path):
    """Property test (hypothesis): the C reader (csrc/host_io.cpp) and the oracle's reader agree on randomly generated
    alignments -- random letters incl. non-standard ones, consistent insert columns ('.' / lowercase), random line wrapping,
    LF / CRLF, blank lines, random max_gap_fraction; errors are raised by both or by neither."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    upper = "ACDEFGHIKLMNPQRSTVWYBJOUXZ-"

    @st.composite
    def alignment(draw):
        ncols = draw(st.integers(1, 40))
        insert = [draw(st.booleans()) and draw(st.booleans()) for _ in range(ncols)]      # ~25 % insert columns
        if all(insert):
            insert[draw(st.integers(0, ncols - 1))] = False
        nseq = draw(st.integers(1, 12))
        recs = []
        for _ in range(nseq):
            chars = []
            for c in range(ncols):
                if insert[c]:
                    chars.append(draw(st.sampled_from("." + "acdefghiklmnpqrstvwy")))
                else:
                    chars.append(draw(st.sampled_from(upper)))
            recs.append("".join(chars))
        eol = draw(st.sampled_from(["\n", "\r\n"]))
        wrap = draw(st.integers(1, 50))
        text = ""
        for k, r in enumerate(recs):
            text += f">s{k} d{eol}"
            for o in range(0, len(r), wrap):
                text += r[o:o + wrap] + eol
            if draw(st.booleans()):
                text += eol
        mg = draw(st.sampled_from([0.0, 0.1, 0.5, 0.9, 1.0]))
        return text, mg

    path = tmp_path / "fuzz.fasta"

    @settings(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
    @given(alignment())
    def run(case):
        text, mg = case
        path.write_bytes(text.encode())
        try:
            want = orc.read_fasta_alignment(str(path), mg)
        except ValueError as e:
            with pytest.raises(ValueError):
                pkg.read_fasta_alignment(str(path), mg)
            assert "passed" in str(e) or "inconsistent" in str(e) or "aligned" in str(e)
            return
        got = pkg.read_fasta_alignment(str(path), mg)
        assert got.dtype == np.int8 and np.array_equal(got, want)
        d_got, keep = pkg.remove_duplicate_sequences(got)
        assert np.array_equal(d_got, orc.remove_duplicate_sequences(want)) and np.array_equal(d_got, got[keep])

    run()


def test_bench_arms_describe_the_same_workload_and_probe_julia():
    """Both arms of bench.py build config.workload from one function (the driver compares the strings), every BASELINE.json config
    has an entry, and the julia probe reports an outcome instead of assuming one."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.workload_string("C") == "synthetic L=500 M=200000 theta=auto score=frob pseudocount=0.8 min_separation=5 (BASELINE.json configs[2])"
    assert {"B", "C", "D", "Cs", "E"} <= set(b.WORKLOADS)
    assert b.WORKLOADS["D"][:4] == (500, 200000, "DI", 0.2) and b.WORKLOADS["E"][:2] == (1500, 1000000)
    p = b.julia_probe()
    assert set(p) == {"julia", "outcome"} and isinstance(p["outcome"], str)


def test_fullsize_oracle_cache_round_trip(orc, tmp_path, monkeypatch):
    """oracle/fullsize.py at a small shape: the staged run equals gdca_from_Z, and a cache hit returns the same results."""
    import importlib
    monkeypatch.setenv("GDCA_ORACLE_CACHE", str(tmp_path))
    from oracle import fullsize
    importlib.reload(fullsize)
    d = fullsize.pipeline_full(40, 1500, "DI", 0.2, seed=5, keep_big=True)
    st = {}
    Ro = orc.gdca_from_Z(orc.synth_alignment(40, 1500, 5), 0.2, "auto", "DI", 5, stages=st)
    assert np.array_equal(d["counts"], st["counts"]) and d["Meff"] == st["Meff"] and d["thresh"] == st["thresh"]
    assert [(int(i), int(j)) for i, j, _ in d["R"].tolist()] == [(i, j) for i, j, _ in Ro]
    assert np.array_equal(d["C"], st["C"]) and not d["cached"]
    d2 = fullsize.pipeline_full(40, 1500, "DI", 0.2, seed=5)
    assert d2["cached"] and d2["weights_cached"] and np.array_equal(d2["R"], d["R"]) and np.array_equal(d2["S"], d["S"])
    assert fullsize.set_all_threads() == (os.cpu_count() or 1) == orc.lib().oracle_max_threads()
