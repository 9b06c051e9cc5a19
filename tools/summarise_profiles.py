"""CPU-side: turn gpurun_out/{launches_bench.csv, top_kernels.ncu-rep, bench_r1.json} into tracked summaries under profiles/."""
import csv, io, json, os, re, subprocess, sys

ROUND = sys.argv[1] if len(sys.argv) > 1 else "r1"
OUT = "profiles"
os.makedirs(OUT, exist_ok=True)

# ---- 1. launch list
if ROUND == "r1":
    agg = subprocess.run([sys.executable, "tools/launch_agg.py", "gpurun_out/launches_bench.csv", "--step", "4"], capture_output=True, text=True).stdout
    how = ("Command (on the GPU box): `ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv "
           "--log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline`; the list below is the "
           "timed step = the 4th hot-path pass (after 3 warm-up passes), cut out with `tools/launch_agg.py --step 4`\n")
else:
    agg = subprocess.run([sys.executable, "tools/launch_agg.py", f"gpurun_out/{ROUND}_launches.csv"], capture_output=True, text=True).stdout
    how = ("Command (on the GPU box, tools/make_profiles_r2.sh): `ncu --profile-from-start off --metrics gpu__time_duration.sum,"
           "launch__grid_size --clock-control none --csv --log-file gpurun_out/r2_launches.csv python tools/profile_step.py C`: ONE "
           "gdca_run_resident step of config C (L=500, M=200k) inside a cudaProfiler range, after two unprofiled warm-up steps\n")
open(f"{OUT}/{ROUND}_launches.md", "w").write(
    f"# {ROUND}: ncu launch list of one hot-path step\n\n" + how +
    "(cold-cache, serialised launches: compare SHARES, not absolutes; the bench value itself is never taken under ncu).\n\n```\n" + agg + "```\n")

# ---- 2. full-set metrics of the top kernels
import glob
reports = sorted(glob.glob("gpurun_out/top_*.ncu-rep" if ROUND == "r1" else f"gpurun_out/{ROUND}_top_*.ncu-rep"))
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
]
seen = set()
traffic = {}
with open(f"{OUT}/{ROUND}_top_kernels.md", "w") as f:
    f.write(f"# {ROUND}: `ncu --set full --clock-control none --import-source on` of the top kernels (config C: L=500, M=200k)\n\n"
            "One launch each, captured with tools/make_profiles.sh; read here with `ncu -i ... --page raw --csv`.\n")
    allrows = []
    for rep in reports:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(io.StringIO(raw)))
        allrows += [(rr[0], rr[1], r) for r in rr[2:]]
    # dgemm: keep the longest launch of each instantiation (the lauum GEMM, the biggest trtri / trailing update)
    best = {}
    for hdr, units, r in allrows:
        name = r[hdr.index("Kernel Name")]
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")
        t = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
        if short not in best or t > best[short][0]:
            best[short] = (t, hdr, units, r)
    for short, (t, hdr, units, r) in best.items():
        if short in seen:
            continue
        seen.add(short)
        f.write(f"\n## {short}\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in KEYS:
            if k in hdr and r[hdr.index(k)] not in ("", "n/a"):
                f.write(f"| `{k}` | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
        def _bytes(key):
            if key not in hdr or not r[hdr.index(key)]:
                return None
            v = float(r[hdr.index(key)].replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(units[hdr.index(key)], 1)
        traffic[short] = {"dram_bytes_read": _bytes("dram__bytes_read.sum"), "dram_bytes_write": _bytes("dram__bytes_write.sum"),
                          "gpu_time_ms_under_ncu": t if units[hdr.index("gpu__time_duration.sum")].startswith("ms") else t / 1e3}
        st = []
        for i, h in enumerate(hdr):
            m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
            if m and r[i]:
                try:
                    st.append((float(r[i].replace(",", "")), m.group(1)))
                except ValueError:
                    pass
        st.sort(reverse=True)
        f.write("\nstall reasons (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in st[:8]) + "\n")

json.dump(traffic, open(f"{OUT}/{ROUND}_traffic.json", "w"), indent=1)

# ---- 3. the bench line of the same build
if os.path.exists(f"gpurun_out/bench_{ROUND}.json"):
    line = open(f"gpurun_out/bench_{ROUND}.json").read().strip().splitlines()[-1]
    d = json.loads(line)
    open(f"{OUT}/{ROUND}_bench.json", "w").write(json.dumps(d, indent=1) + "\n")
print(open(f"{OUT}/{ROUND}_top_kernels.md").read()[:3000])
