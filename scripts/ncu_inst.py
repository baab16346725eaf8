"""Warp-instruction budget per source line (from ncu --page source --csv --print-source cuda,sass)."""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
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
        items.append((ie, cur, r[0], r[1].strip()[:80], s, te))
        tot += ie
print("total warp instructions", tot)
for ie, f, ln, src, s, te in sorted(items, reverse=True)[:top]:
    print("%5.1f%% %11d thr/inst %5.1f samples %6d %-13s:%-4s %s" % (100.0 * ie / tot, ie, te / max(ie, 1), s, f, ln, src))
