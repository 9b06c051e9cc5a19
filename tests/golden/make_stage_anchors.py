"""Writes tests/golden/stage_anchors.json from the CPU oracle (secondary anchors)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import gdca_oracle as o  # noqa: E402

cases = {
    "small_default": ("small.fasta.gz", dict()),
    "small_DI_dedup": ("small.fasta.gz", dict(pseudocount=0.2, score="DI", remove_dups=True)),
    "small_DI_theta0": ("small.fasta.gz", dict(pseudocount=0.2, score="DI", theta=0.0, max_gap_fraction=0.8, min_separation=4)),
    "large_DI_dedup": ("large.fasta.gz", dict(pseudocount=0.2, score="DI", remove_dups=True)),
}
out = {}
for name, (fa, kw) in cases.items():
    st = {}
    R = o.gDCA(os.path.join(HERE, fa), stages=st, **kw)
    vals, h = np.unique(st["counts"], return_counts=True)
    C, mJ = st["C"], st["mJ"]
    out[name] = dict(
        M=int(st["W"].shape[0]), L=int(C.shape[0] // (st["q"] - 1)), q=int(st["q"]), theta=st["theta"],
        thresh=int(st["thresh"]), Meff=st["Meff"], count_hist={int(v): int(c) for v, c in zip(vals, h)},
        C00=C[0, 0], C01=C[0, 1], C0_20=C[0, 20], traceC=float(np.trace(C)), mJ00=mJ[0, 0], trace_mJ=float(np.trace(mJ)),
        top=[R[0][0], R[0][1], R[0][2]], nrows=len(R))
json.dump(out, open(os.path.join(HERE, "stage_anchors.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1)[:600])
