"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump over source line ranges of one file.
usage: python tools/ncu_phases.py dump.csv file name:lo-hi [name:lo-hi ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
fname = sys.argv[2]
ranges = []
for a in sys.argv[3:]:
    n, r = a.split(":"); lo, hi = r.split("-"); ranges.append((n, int(lo), int(hi)))
cur = None
tot_s = tot_i = 0
agg = {n: [0, 0] for n, _, _ in ranges}
agg["other"] = [0, 0]
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if not r or not r[0].strip().isdigit() or len(r) < 8:
        continue
    try:
        ln = int(r[0]); s = int(r[4] or 0); i = int(r[7] or 0)
    except ValueError:
        continue
    tot_s += s; tot_i += i
    key = "other"
    if cur == fname:
        for n, lo, hi in ranges:
            if lo <= ln <= hi:
                key = n; break
    agg[key][0] += s; agg[key][1] += i
for k, (s, i) in agg.items():
    print("%-12s samples %6.2f%%  warp-inst %6.2f%%  (%d inst)" % (k, 100.0 * s / max(tot_s, 1), 100.0 * i / max(tot_i, 1), i))
