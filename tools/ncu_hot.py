"""Hot SASS regions of one profiled kernel: `ncu -i rep --page source --print-source sass --csv` -> runs of instructions
executed about once per iteration of the hottest loop, with instructions per iteration and stall samples.
usage: python tools/ncu_hot.py report.ncu-rep [marker-opcode (default CREDUX)] [dump-file]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "CREDUX"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))[2:]
ops = [(int(r[5] or 0), r[1].strip(), int(r[4] or 0), float(r[8] or 0)) for r in rows if len(r) > 8]
tot = sum(o[0] for o in ops)
tot_s = sum(o[2] for o in ops)
it = max([o[0] for o in ops if marker in o[1]] or [1])
print("warp-instructions %.3f G, samples %d, iterations of the loop holding %s: %d" % (tot / 1e9, tot_s, marker, it))
runs, cur = [], None
for i, o in enumerate(ops):
    if o[0] > 0.25 * it:
        if cur is None:
            cur = [i, i]
        cur[1] = i
    elif cur is not None and i - cur[1] > 30:
        runs.append(cur); cur = None
if cur:
    runs.append(cur)
for a, b in runs:
    t = sum(o[0] for o in ops[a:b + 1]); sm = sum(o[2] for o in ops[a:b + 1])
    h = collections.Counter()
    for o in ops[a:b + 1]:
        w = o[1].split()
        h[(w[1] if w[0].startswith("@") else w[0]).split(".")[0]] += o[0]
    print("sass %5d-%5d  inst/iter %6.1f  inst %5.1f%%  samples %5.1f%%  %s" % (a, b, t / it, 100.0 * t / tot, 100.0 * sm / tot_s,
          " ".join("%s:%.0f" % (k, v / it) for k, v in h.most_common(8))))
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        for i, o in enumerate(ops):
            f.write("%5d %6.2f %6d %3.0f %s\n" % (i, o[0] / it, o[2], o[3], o[1][:110]))
