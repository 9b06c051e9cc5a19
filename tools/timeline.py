"""Kernel timeline of ONE hot-path step (CUPTI through torch.profiler: start, duration and stream of every launch, with the real
overlap between streams -- what the serialised ncu launch list cannot show).
   python tools/timeline.py [C|D|B] [out.json]   -> gpurun_out/timeline_<name>.json: [[name, stream, start_us, dur_us], ...]"""
import ctypes, json, sys, os
sys.path.insert(0, ".")
import torch
import __graft_entry__ as g
pkg = g.load_package()
from gaussdca_jl_b200 import _lib as glib
name = sys.argv[1] if len(sys.argv) > 1 else "C"
out = sys.argv[2] if len(sys.argv) > 2 else f"gpurun_out/timeline_{name}.json"
L, M, score, pc = {"B": (200, 50000, "frob", 0.8), "C": (500, 200000, "frob", 0.8), "D": (500, 200000, "DI", 0.2)}[name]
ctx = pkg.Context(0)
lib = ctx.lib
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
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
tmp = out + ".trace.json"
prof.export_chrome_trace(tmp)
tr = json.load(open(tmp))
ev = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"] if ev else 0
rows = [[e["name"][:60], e.get("args", {}).get("stream"), round(e["ts"] - t0, 2), round(e["dur"], 2)] for e in ev]
json.dump(rows, open(out, "w"))
os.remove(tmp)
print({k: round(v, 3) for k, v in st.asdict().items() if k.startswith("ms_")}, len(rows), "events ->", out)
