"""Record the pipe peaks gdca_probe_peaks measures on this GPU (LOP3, POPC, DMMA, DFMA) as a file: the roofline denominators
bench.py uses for the INT32 and FP64 kernels (MEASURED_PEAKS.json carries HBM and bf16 only).  Run under gpurun; copy the
output to profiles/r2_probe_peaks.json."""
import ctypes, json, subprocess, sys
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package()
ctx = pkg.Context(0)
runs = []
for _ in range(3):
    v = [ctypes.c_double() for _ in range(4)]
    ctx.check(ctx.lib.gdca_probe_peaks(ctx.h, *[ctypes.byref(x) for x in v]))
    runs.append([x.value for x in v])
best = [max(r[i] for r in runs) for i in range(4)]
smi = subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader"],
                     capture_output=True, text=True).stdout.strip()
print(json.dumps({"lop3_tops": best[0], "popc_tops": best[1], "dmma_tflops": best[2], "dfma_tflops": best[3], "runs": runs,
                  "how": "gdca_probe_peaks (csrc/probe.cu): register-resident dependent-chain microkernels, 148 x 8 CTAs, best of 3",
                  "nvidia_smi": smi}, indent=1))
