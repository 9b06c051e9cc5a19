"""Print the key metrics of an .ncu-rep (read here on the CPU box):  python tools/ncu_summary.py rep [regex]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if pat and not pat.search(name):
        continue
    print("=====", name[:100])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:80s} {r[i]:>18s} {units[i]}")
    # stall reasons, sorted
    st = []
    for i, h in enumerate(hdr):
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)
        if m and r[i]:
            try:
                st.append((float(r[i].replace(",", "")), m.group(1)))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("  stalls (warps per issue-active):", ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))
