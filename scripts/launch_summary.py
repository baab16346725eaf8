"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): python scripts/launch_summary.py launches.csv "command" """
import csv, sys, collections, re
path = sys.argv[1]
cmd = sys.argv[2] if len(sys.argv) > 2 else ""
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows:
    if r is hdr or len(r) <= vi or r[ki] == "Kernel Name":
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    u = r[ui]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    name = re.sub(r"\(.*$", "", r[ki]).replace("void ", "").replace("vr::", "").strip()
    c, t = tot.get(name, (0, 0.0))
    tot[name] = (c + 1, t + v)
allus = sum(t for _, t in tot.values())
print("ncu launch list of:", cmd)
print("(--metrics gpu__time_duration.sum --clock-control none -c 600; cold-cache, serialised: compare SHARES)")
print("%-60s %6s %12s %7s" % ("kernel", "count", "total_us", "share"))
for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %6d %12.1f %6.1f%%" % (name[:60], c, t, 100 * t / allus))
print("%-60s %6d %12.1f" % ("all", sum(c for c, _ in tot.values()), allus))
