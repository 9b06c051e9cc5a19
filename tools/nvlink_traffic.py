"""NVLink bytes moved per gDCA step on a device group: nvidia-smi per-link data counters read before and after K steps.
    python tools/nvlink_traffic.py [ngpu] [C|E] [steps]        (run under gpurun --gpus N; output -> profiles/r2_nvlink.json)
Expected at config C, N GPUs: alignment broadcast 100 MB per receiving member; covariance rows stored
into the leader ~ (N-1)/N x 400 MB (upper site blocks); factor panels + inverted diagonal blocks pulled by every member ~ 410 MB
each; X21 column slices stored to every member ~ (N-1)/N x 400 MB out of each member's share; lauum row tiles stored to the
leader ~ (N-1)/N x 400 MB; neighbour hits (peer atomics) negligible."""
import ctypes, json, re, subprocess, sys
sys.path.insert(0, ".")
import torch
import __graft_entry__ as g
pkg = g.load_package()
from gaussdca_jl_b200 import _lib as glib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
name = sys.argv[2] if len(sys.argv) > 2 else "C"
K = int(sys.argv[3]) if len(sys.argv) > 3 else 5
L, M, score, pc = {"C": (500, 200000, "frob", 0.8), "E": (1500, 1000000, "frob", 0.8)}[name]


def counters():
    out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d"], capture_output=True, text=True).stdout
    res, gpu = {}, None
    for line in out.splitlines():
        m = re.match(r"GPU (\d+):", line)
        if m:
            gpu = int(m.group(1))
            res[gpu] = {"tx_kib": 0, "rx_kib": 0}
            continue
        m = re.search(r"Data (Tx|Rx): (\d+) KiB", line)
        if m and gpu is not None:
            res[gpu]["tx_kib" if m.group(1) == "Tx" else "rx_kib"] += int(m.group(2))
    return res


ctx = pkg.Context(devices=list(range(N)))
lib = ctx.lib
Zd = torch.empty((M, L), dtype=torch.int8, device="cuda:0")
ctx.check(lib.gdca_synth_alignment_dev(ctx.h, ctypes.c_void_p(Zd.data_ptr()), L, M, 20140321))
n_out = int(lib.gdca_ranking_length(L, 5))
st = glib.Stats()
def step():
    ctx.check(lib.gdca_run_resident(ctx.h, ctypes.c_void_p(Zd.data_ptr()), L, M, -1.0, pc, glib.SCORE_CODES[score], 5, None, n_out, ctypes.byref(st)))
for _ in range(2):
    step()
for d in range(N):
    torch.cuda.synchronize(d)
c0 = counters()
for _ in range(K):
    step()
for d in range(N):
    torch.cuda.synchronize(d)
c1 = counters()
per = {str(gpu): {k: (c1[gpu][k] - c0[gpu][k]) * 1024 / K for k in c1[gpu]} for gpu in sorted(c1) if gpu < N}
print(json.dumps({"workload": name, "n_gpus": N, "steps": K, "ms_per_step": st.ms_total,
                  "nvlink_bytes_per_step_per_gpu": per,
                  "total_tx_bytes_per_step": sum(v["tx_kib"] for v in per.values()),
                  "how": "nvidia-smi nvlink -gt d (sum over the links of each GPU), before / after K gdca_run_resident steps on a gdca_create_multi group",
                  "shared_factorisation": int(lib.gdca_dev_inverse_shared(ctx.h))}, indent=1))
