"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck): fixtures + ragged synthetic shapes."""
import sys
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
orc = g.load_oracle()
ctx = pkg.Context(0)
# force the tcgen05 / TMA / TMEM prefilter on (it switches itself on only for M >= 16384): the sanitized path is then the
# whole production path -- prefilter -> exact sweep of the flagged blocks -> covariance -> inverse -> scores -> ranking
import os
if os.environ.get("GDCA_SANITIZE_FILTER", "1") != "0":
    ctx.check(ctx.lib.gdca_set_tc_filter(ctx.h, 2))
# round 2: the covariance on the tensor cores (cta_group::2 pairs) whenever the weights are count classes, the candidate-pair
# exact stage (default), the blocked diagonal kernel, and one shape that takes the sliced INT8 inversion (n = 2560) -- which is
# also the CUDA-graph replay path on the second call
ctx.set_cov_engine(2)
R = pkg.gDCA("tests/golden/small.fasta.gz", ctx=ctx)
print("small frob", R[0])
R = pkg.gDCA("tests/golden/small.fasta.gz", pseudocount=0.2, score="DI", remove_dups=True, ctx=ctx)
print("small DI", R[0])
big = [(128, 700)] if os.environ.get("GDCA_SANITIZE_BIG", "1") != "0" else []
for L, M in [(33, 129), (70, 300), (130, 257)] + big + big:
    Z = orc.synth_alignment(L, M, seed=L)
    for score in ("frob", "DI"):
        R = pkg.gdca_from_alignment(Z, score=score, ctx=ctx)
        print(L, M, score, R[0])
