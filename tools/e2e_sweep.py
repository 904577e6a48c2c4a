"""End-to-end step (th_host_run on host buffers) over chunk sizes and lane counts, one batch generated once.
usage: python tools/e2e_sweep.py n_reads "chunk:lanes,chunk:lanes,..." [reps]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidehunter_b200 as T
from tidehunter_b200 import synth

n = int(sys.argv[1])
cfgs = [tuple(int(x) for x in c.split(":")) for c in sys.argv[2].split(",")]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
names, seqs = synth.gen_reads("r2c2", n)
batch = T.Batch(names, seqs)
for chunk, lanes in cfgs:
    th = T.TideHunter(out_fmt=1, chunk_reads=chunk, lanes=lanes)
    ts = []
    for rep in range(reps + 1):
        t0 = time.perf_counter()
        out = th.run(batch, copy=False)
        ts.append(1e3 * (time.perf_counter() - t0))
    print("chunk %5d lanes %d: ms per step %s -> best %.0f reads/s" % (chunk, lanes, " ".join("%.1f" % t for t in ts[1:]), n / (min(ts[1:]) * 1e-3)), flush=True)
    th.close()
