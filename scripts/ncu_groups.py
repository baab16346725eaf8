"""Warp-stall samples and instructions of an ncu source page grouped by function (line ranges of
vr_device.cuh / vr_trace.cu given as NAME:FILE:FIRST-LAST ...).
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python scripts/ncu_groups.py NAME:FILE:A-B ..."""
import csv, sys
groups = []
for g in sys.argv[1:]:
    name, f, rng = g.split(":")
    a, b = rng.split("-")
    groups.append((name, f, int(a), int(b)))
rows = list(csv.reader(sys.stdin))
cur = None
acc = {}
tot = [0, 0, 0]
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 10 and r[0] not in ("", "Line No"):
        try:
            ln = int(r[0]); s = int(r[4]); ie = int(r[7]); te = int(r[8])
        except ValueError:
            continue
        key = "other:" + str(cur)
        for name, f, a, b in groups:
            if cur == f and a <= ln <= b:
                key = name
                break
        v = acc.setdefault(key, [0, 0, 0])
        v[0] += s; v[1] += ie; v[2] += te
        tot[0] += s; tot[1] += ie; tot[2] += te
print("%-28s %8s %6s %12s %6s %8s" % ("group", "samples", "%", "warp inst", "%", "thr/inst"))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %8d %5.1f%% %12d %5.1f%% %8.1f" % (k, v[0], 100.0 * v[0] / max(tot[0], 1), v[1], 100.0 * v[1] / max(tot[1], 1), v[2] / max(v[1], 1)))
print("%-28s %8d        %12d        %8.1f" % ("total", tot[0], tot[1], tot[2] / max(tot[1], 1)))
