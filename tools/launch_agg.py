"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = collections.OrderedDict(), 0.0
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
    v = float(r[vi].replace(",", ""))
    v = {"ns": v / 1e6, "us": v / 1e3, "usecond": v / 1e3, "ms": v, "msecond": v, "s": v * 1e3}.get(r[ui], v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:58]:58s} {c:8d} {v:10.3f} {100*v/tot:6.2f}% {1e3*v/c:9.1f}")
print(f"{'TOTAL':58s} {sum(c for c,_ in agg.values()):8d} {tot:10.3f}")
