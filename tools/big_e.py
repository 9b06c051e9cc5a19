"""Config E (L=1500, M=1M) functional run on ONE GPU: checks sizes/overflow paths; prints stage times."""
import ctypes, json, sys, time
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package()
from gaussdca_jl_b200 import _lib as glib
import torch
L, M = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1500, 1_000_000)
ctx = pkg.Context(0)
lib = ctx.lib
Zd = torch.empty((M, L), dtype=torch.int8, device="cuda:0")
ctx.check(lib.gdca_synth_alignment_dev(ctx.h, ctypes.c_void_p(Zd.data_ptr()), L, M, 20140321))
n_out = int(lib.gdca_ranking_length(L, 5))
R = np.empty(n_out, dtype=glib.RANK_DTYPE)
st = glib.Stats()
t = time.time()
ctx.check(lib.gdca_run_resident(ctx.h, ctypes.c_void_p(Zd.data_ptr()), L, M, -1.0, 0.8, 0, 5, glib.ptr(R), n_out, ctypes.byref(st)))
d = st.asdict(); d["wall_s"] = time.time() - t
d["top"] = [int(R["i"][0]), int(R["j"][0]), float(R["score"][0])]
d["sorted"] = bool(np.all(np.diff(R["score"]) <= 0)); d["rows"] = int(n_out)
d["mem_GB"] = torch.cuda.mem_get_info()[0] / 1e9
print(json.dumps(d))
