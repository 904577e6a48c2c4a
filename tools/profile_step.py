"""One device-resident step of the hot path over N synthetic R2C2 reads: the command ncu wraps.
usage: python tools/profile_step.py [n_reads] [steps] [shape]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidehunter_b200 as T
from tidehunter_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
shape = sys.argv[3] if len(sys.argv) > 3 else "r2c2"
names, seqs = synth.gen_reads(shape, n)
ctx = T.GpuContext()
ctx.upload(seqs)
hist = []
for _ in range(steps):
    r = ctx.process_resident()
    hist.append(r.stats.as_dict())
print({k: v for k, v in r.stats.as_dict().items()})
if steps > 2:  # the POA kernel's time varies from launch to launch: all steps but the first, sorted
    for key in ("ms_poa", "ms_ksw", "ms_chain", "ms_total"):
        print(key, "per step:", " ".join("%.1f" % h[key] for h in sorted(hist[1:], key=lambda h: h[key])))
    print("ms_poa in launch order:", " ".join("%.1f" % h["ms_poa"] for h in hist))
cn = ctx.counters()
ph = cn[16:23]
names = ["setup", "rows", "backtrack", "merge", "reorder", "consensus", "total"]
print("poa phase share of warp-cycles:", {k: round(v / max(ph[6], 1), 4) for k, v in zip(names, ph)})
print("poa warp-cycles per task: %.0f" % (ph[6] / max(r.stats.n_tasks, 1)))
ctx.close()
