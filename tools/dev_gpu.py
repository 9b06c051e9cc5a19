"""Developer scratch run on the GPU box: pipe probes + stage timings at the BASELINE shapes."""
import ctypes
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as g

pkg = g.load_package()
from gaussdca_jl_b200._lib import ptr  # noqa: E402

ctx = pkg.Context(0)
lop3, popc, dmma, dfma = (ctypes.c_double() for _ in range(4))
ctx.check(ctx.lib.gdca_probe_peaks(ctx.h, ctypes.byref(lop3), ctypes.byref(popc), ctypes.byref(dmma), ctypes.byref(dfma)))
print(json.dumps(dict(lop3_tops=lop3.value, popc_tops=popc.value, dmma_tflops=dmma.value, dfma_tflops=dfma.value)))

shapes = [(200, 50000), (500, 200000)] if len(sys.argv) < 2 else [tuple(map(int, a.split("x"))) for a in sys.argv[1:]]
for L, M in shapes:
    Z = np.empty((M, L), dtype=np.int8)
    ctx.check(ctx.lib.gdca_synth_alignment(ctx.h, ptr(Z), L, M, 20140321))
    for score in ("frob", "DI"):
        for rep in range(2):
            t = time.time()
            R, st = pkg.gdca_from_alignment(Z, score=score, pseudocount=0.8 if score == "frob" else 0.2, ctx=ctx,
                                            return_stats=True, as_array=True)
            wall = time.time() - t
        st["wall_s"] = wall
        st["score"] = score
        st["top"] = [int(R["i"][0]), int(R["j"][0]), float(R["score"][0])]
        print(json.dumps(st))
        sys.stdout.flush()
