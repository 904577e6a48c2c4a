"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: python tools/ncu_lines.py dump.csv [top_n] [file_filter]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
flt = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
cur_file = None
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if not r or not r[0].strip().isdigit() or len(r) < 8:
        continue
    try:
        ln = int(r[0]); samples = int(r[4] or 0); inst = int(r[7] or 0)
    except ValueError:
        continue
    a = agg.setdefault((cur_file, ln), [0, 0, r[1]])
    a[0] += samples
    a[1] += inst
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print("total samples", tot_s, "total warp-inst", tot_i)
items = [(k, a) for k, a in agg.items() if not flt or flt in k[0]]
for key, name in ((0, "samples"), (1, "instructions")):
    print("--- top by", name)
    for k, a in sorted(items, key=lambda x: -x[1][key])[:top]:
        print("%-16s %4d  s=%5.2f%% i=%5.2f%%  %s" % (k[0], k[1], 100 * a[0] / tot_s, 100 * a[1] / tot_i, a[2].strip()[:120]))
