"""Append the `ncu --set full` summary of one report to profiles/<round>_top_kernels.md and <round>_traffic.json:
   python tools/append_profile.py gpurun_out/r2_di_eig.ncu-rep r2 "note printed under the heading" """
import csv, io, json, re, subprocess, sys

sys.argv += [""] * 3
rep, rnd, note = sys.argv[1], sys.argv[2] or "r2", sys.argv[3]
src = open("tools/summarise_profiles.py").read()
KEYS = eval(re.search(r"KEYS = (\[.*?\n\])", src, re.S).group(1))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]
traffic = json.load(open(f"profiles/{rnd}_traffic.json"))
with open(f"profiles/{rnd}_top_kernels.md", "a") as f:
    for r in rr[2:]:
        name = r[hdr.index("Kernel Name")]
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")
        f.write(f"\n## {short}\n\n" + (note + "\n\n" if note else "") + "| metric | value | unit |\n|---|---|---|\n")
        for k in KEYS:
            if k in hdr and r[hdr.index(k)] not in ("", "n/a"):
                f.write(f"| `{k}` | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
        def _bytes(key):
            v = float(r[hdr.index(key)].replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(units[hdr.index(key)], 1)
        t = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
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
        print(short, traffic[short])
json.dump(traffic, open(f"profiles/{rnd}_traffic.json", "w"), indent=1)
