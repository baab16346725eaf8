"""Top source lines by warp-stall samples from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass [-k id] | python scripts/ncu_lines.py [N]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rows = list(csv.reader(sys.stdin))
cur = None
items = []
tot = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 10 and r[0] not in ("", "Line No"):
        try:
            s = int(r[4]); ie = int(r[7]); te = int(r[8])
        except ValueError:
            continue
        items.append((s, cur, r[0], r[1].strip()[:90], ie, te))
        tot += s
print("total samples", tot)
for s, f, ln, src, ie, te in sorted(items, reverse=True)[:top]:
    print("%6d %5.1f%%  %-14s:%-4s inst %9d thr/inst %5.1f  %s" % (s, 100.0 * s / max(tot, 1), f, ln, ie, te / max(ie, 1), src))
