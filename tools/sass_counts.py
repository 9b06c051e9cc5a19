"""cuobjdump -sass of the in-tree library: per kernel, how often the instructions that prove the Blackwell-native paths occur
(tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG, FP64 tensor -> DMMA, cp.async -> LDGSTS).
    python tools/sass_counts.py > profiles/r2_sass.txt        (CPU box; no GPU needed)"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "gaussdca.jl_b200/libgdca_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
PAT = ["UTCIMMA", "UTCQMMA", "UTCOMMA", "UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCATOMSWS", "DMMA",
       "HMMA", "IMMA", "LDGSTS", "SYNCS", "ELECT", "LOP3", "POPC", "REDG", "ATOMG", "RED.E", "ATOM.E"]
cur, counts, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("void ", "")
        cur = counts.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cur["_total"] += 1
        for p in PAT:
            if op.startswith(p):
                cur[p + ("." + ".".join(op.split(".")[1:3]) if p in ("UTMALDG", "DMMA", "UTCBAR") and "." in op else "")] += 1
print(f"# cuobjdump -sass {so}: instruction counts per kernel (architectures in the fatbin: {', '.join(sorted(arch))})")
print("# tcgen05.mma = UTC*MMA (UTCIMMA int8, UTCQMMA fp8, UTCOMMA mxf4), tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG, FP64 tensor = DMMA\n")
for name, c in counts.items():
    hot = {k: v for k, v in c.items() if k != "_total" and not k.startswith(("LOP3", "POPC", "SYNCS", "ELECT", "REDG", "ATOMG", "RED.E", "ATOM.E"))}
    misc = {k: v for k, v in c.items() if k.startswith(("LOP3", "POPC", "REDG", "ATOMG", "RED.E", "ATOM.E"))}
    print(f"{name[:90]:90s} total {c['_total']:6d}  " + "  ".join(f"{k} {v}" for k, v in sorted(hot.items())) +
          ("   | " + "  ".join(f"{k} {v}" for k, v in sorted(misc.items())) if misc else ""))
