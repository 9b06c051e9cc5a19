import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def orc():
    import __graft_entry__ as g
    o = g.load_oracle()
    o.build()
    return o


@pytest.fixture(scope="session")
def ctx(pkg):
    # no skip: on a GPU box the CUDA library must load and run, or the test fails loudly
    return pkg.Context(0)


def golden_path(name):
    return os.path.join(GOLDEN, name)


def read_golden(name):
    import gzip
    d = {}
    with gzip.open(golden_path(name + ".txt.gz"), "rt") as fh:
        for line in fh:
            sl = line.split()
            if not sl:
                continue
            assert len(sl) == 3
            k = (int(sl[0]), int(sl[1]))
            assert k not in d
            d[k] = float(sl[2])
    return d


# the reference's own test matrix (test/runtests.jl:52-76): golden file, fasta, kwargs
GOLDEN_CASES = [
    ("small.FNRout", "small.fasta.gz", dict()),
    ("small.DIRout", "small.fasta.gz", dict(pseudocount=0.2, score="DI", remove_dups=True)),
    ("small.DIRout2", "small.fasta.gz", dict(pseudocount=0.2, score="DI", theta=0.0, max_gap_fraction=0.8,
                                              min_separation=4)),
    ("large.DIRout", "large.fasta.gz", dict(pseudocount=0.2, score="DI", remove_dups=True)),
]


def rank_todict(R):
    d = {}
    for i, j, x in R:
        assert (i, j) not in d
        d[(i, j)] = x
    return d


def printed_todict(text):
    """todict of test/runtests.jl:29-39"""
    d = {}
    for line in text.splitlines():
        sl = line.split()
        if not sl:
            continue
        assert len(sl) == 3
        k = (int(sl[0]), int(sl[1]))
        assert k not in d
        d[k] = float(sl[2])
    return d
