"""Where the end-to-end step goes: Python marshalling, th_host_run (set TH_HOST_TIMING=1 for its own breakdown), copy-out.
usage: python tools/e2e_probe.py [n_reads] [chunk] [lanes] [reps]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidehunter_b200 as T
from tidehunter_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24576
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 3
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
names, seqs = synth.gen_reads("r2c2", n)
th = T.TideHunter(out_fmt=1, chunk_reads=chunk, lanes=lanes)
h = T.host_lib()
for rep in range(reps + 2):
    t0 = time.perf_counter()
    bs, seqs_a, lens_a = T._arrays(seqs)
    names_a = (C.c_char_p * len(names))(*names)
    t1 = time.perf_counter()
    out_len = C.c_size_t(0)
    ptr = h.th_host_run(th._h, len(bs), names_a, seqs_a, lens_a, C.byref(out_len))
    t2 = time.perf_counter()
    txt = C.string_at(ptr, out_len.value)
    t3 = time.perf_counter()
    print("rep %d: marshal %.1f ms, th_host_run %.1f ms, copy-out %.1f ms (%d bytes) -> %.0f reads/s" %
          (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), len(txt), n / (t3 - t0)), flush=True)
th.close()
