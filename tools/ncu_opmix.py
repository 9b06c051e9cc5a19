"""Executed-instruction mix by opcode from the source page of an .ncu-rep:  python tools/ncu_opmix.py rep kernel_regex"""
import csv, io, re, subprocess, sys
from collections import Counter
rep, pat = sys.argv[1], re.compile(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
cur, hdr = None, None
mix, stall = Counter(), Counter()
tot = 0
done = set()
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "Kernel Name":
        cur = row[1]
        if cur in done: cur = None   # only the first launch of each kernel
        continue
    if row[0] == "Address":
        hdr = row; continue
    if cur is None or not pat.search(cur): continue
    done_key = cur
    op = row[1].strip().split()[0]
    if op.startswith("@"): op = row[1].strip().split()[1]
    op = op.split(".")[0] if not op.startswith("IMAD") else ".".join(op.split(".")[:2]) if op.startswith("IMAD.MOV") or op.startswith("IMAD.IADD") else "IMAD"
    n = int(row[hdr.index("Instructions Executed")] or 0)
    mix[op] += n; tot += n
    stall[op] += int(row[hdr.index("# Samples")] or 0)
print("total warp-instructions", tot)
for op, n in mix.most_common(18):
    print(f"{op:12s} {n:14d} {100*n/tot:6.2f}%   samples {stall[op]}")
