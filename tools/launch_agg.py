"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections, csv, re, sys
"""   python tools/launch_agg.py launches.csv [--step N]
--step N keeps only the launches of the N-th hot-path step (1-based): from the N-th maxq_kernel (first kernel of a step)
up to the next one, without the L2-flush fill between steps."""
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
if "Metric Name" in hdr:  # lists captured with more than one metric: keep the durations
    mi = hdr.index("Metric Name")
    rows = [hdr] + [r for r in rows[1:] if "gpu__time_duration" in r[mi]]
if "--step" in sys.argv:
    n = int(sys.argv[sys.argv.index("--step") + 1])
    starts = [i for i, r in enumerate(rows) if i and "maxq_kernel" in r[ki]]
    lo = starts[n - 1]
    hi = starts[n] if n < len(starts) else len(rows)
    body = rows[lo:hi]
    while body and ("FillFunctor" in body[-1][ki] or "elementwise" in body[-1][ki]):
        body.pop()
    rows = [hdr] + body
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
