"""Experiment: K contexts on K host threads, each with its own resident sub-batch, vs one context with all reads.
usage: python tools/overlap_test.py n_reads K"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidehunter_b200 as T
from tidehunter_b200 import synth
n = int(sys.argv[1]); K = int(sys.argv[2])
names, seqs = synth.gen_reads("r2c2", n)
per = (n + K - 1) // K
ctxs = [T.GpuContext() for _ in range(K)]
for k, c in enumerate(ctxs):
    c.upload(seqs[k * per:(k + 1) * per])
def run(c, reps):
    for _ in range(reps):
        c.process_resident()
for c in ctxs: run(c, 1)
t0 = time.perf_counter()
th = [threading.Thread(target=run, args=(c, 3)) for c in ctxs]
for t in th: t.start()
for t in th: t.join()
dt = (time.perf_counter() - t0) / 3
print("K=%d share=%s: %.1f ms per %d reads -> %.0f reads/s" % (K, os.environ.get("TH_GPU_SHARE", "1"), dt * 1e3, n, n / dt))
