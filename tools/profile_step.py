"""One hot-path step inside a cudaProfiler range (ncu --profile-from-start off): warm-up steps run unprofiled.
   python tools/profile_step.py [C|D|B] [ozaki 0|1]"""
import ctypes, sys
sys.path.insert(0, ".")
import torch
import __graft_entry__ as g
pkg = g.load_package()
from gaussdca_jl_b200 import _lib as glib
name = sys.argv[1] if len(sys.argv) > 1 else "C"
L, M, score, pc = {"B": (200, 50000, "frob", 0.8), "C": (500, 200000, "frob", 0.8), "D": (500, 200000, "DI", 0.2)}[name]
ctx = pkg.Context(0)
lib = ctx.lib
if len(sys.argv) > 2:
    ctx.check(lib.gdca_set_ozaki(ctx.h, int(sys.argv[2])))
Zd = torch.empty((M, L), dtype=torch.int8, device="cuda:0")
ctx.check(lib.gdca_synth_alignment_dev(ctx.h, ctypes.c_void_p(Zd.data_ptr()), L, M, 20140321))
n_out = int(lib.gdca_ranking_length(L, 5))
st = glib.Stats()
def step():
    ctx.check(lib.gdca_run_resident(ctx.h, ctypes.c_void_p(Zd.data_ptr()), L, M, -1.0, pc, glib.SCORE_CODES[score], 5, None, n_out,
                                    ctypes.byref(st)))
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print({k: round(v, 3) for k, v in st.asdict().items() if k.startswith("ms_")})
