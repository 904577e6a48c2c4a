#!/usr/bin/env python
"""End-to-end throughput (host buffers in, output text out, through th_host_run over the C ABI) for every BASELINE.json
config shape, next to the unmodified reference on a sample of the same reads (all host cores).  One JSON per run:
  python tools/config_bench.py [scale] -> gpurun_out/config_bench.json
bench.py stays the contract's line for configs[1]; this is the table for the other configs."""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    sc = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    import oracle_py as O
    import tidehunter_b200 as T
    from tidehunter_b200 import synth
    with gzip.open(os.path.join(ROOT, "tests", "golden", "golden.json.gz"), "rt") as f:
        ad = json.load(f)["adapters"]
    five, three = ad["five"], ad["three"]
    cores = os.cpu_count() or 1
    cfgs = [
        ("configs[1] r2c2 10 kb, -f 1", lambda n: synth.gen_reads("r2c2", n, start=600000), 32768, ["-f", "1"], dict(out_fmt=1)),
        ("configs[2] short units 50-200 bp x 20-50, -f 2", lambda n: synth.gen_reads("short", n, start=600000), 65536, ["-f", "2"], dict(out_fmt=2)),
        ("configs[3] long units 4-5 kb x 2-4, -f 2", lambda n: synth.gen_reads("long", n, start=600000), 8192, ["-f", "2"], dict(out_fmt=2)),
        ("configs[4] adapters -5 -3 -u -f 2", lambda n: synth.gen_reads("r2c2", n, start=600000, adapters=(five, three)), 16384, ["ADAPTERS", "-u", "-f", "2"],
         dict(out_fmt=2, five_seq=five, three_seq=three, only_unit=1)),
        ("configs[4] adapters -5 -3 -F -f 2 (full-length consensus)", lambda n: synth.gen_reads("r2c2", n, start=600000, adapters=(five, three), three_rc=True), 16384, ["ADAPTERS", "-F", "-f", "2"],
         dict(out_fmt=2, five_seq=five, three_seq=three, only_full_length=1)),
    ]
    rows = []
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        p5, p3 = os.path.join(td, "5.fa"), os.path.join(td, "3.fa")
        open(p5, "w").write(">5\n%s\n" % five)
        open(p3, "w").write(">3\n%s\n" % three)
        for tag, gen, n, argv, kw in cfgs:
            n = max(int(n * sc), 64)
            names, seqs = gen(n)
            bases = synth.total_bases(seqs)
            th = T.TideHunter(device=0, **kw)
            th.run(names, seqs)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                out = th.run(names, seqs)
                ts.append(time.perf_counter() - t0)
            st = th.stats()
            th.close()
            dt = sorted(ts)[1]
            ns = max(min(n, int(1500 * sc) if "long" in tag else int(4096 * sc)), 32)       # reference sample
            path = os.path.join(td, "s.fa")
            O.write_fasta(path, names[:ns], seqs[:ns])
            argv2 = sum((["-5", p5, "-3", p3] if a == "ADAPTERS" else [a] for a in argv), [])
            t0 = time.perf_counter()
            ref = subprocess.run([O.REF_BIN, "-t", str(cores)] + argv2 + [path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            t_ref = time.perf_counter() - t0
            th2 = T.TideHunter(device=0, **kw)
            ours_s = th2.run(names[:ns], seqs[:ns])
            th2.close()
            row = {"config": tag, "reads": n, "bases": bases, "e2e_reads_per_s": round(n / dt, 1), "e2e_gbp_per_s": round(bases / dt / 1e9, 4), "e2e_ms": round(1e3 * dt, 1),
                   "device_ms_sum_over_chunks": round(st["ms_total"], 1), "records_bytes": len(out),
                   "reference": {"reads": ns, "reads_per_s": round(ns / t_ref, 1), "cores": cores, "identical_on_sample": ours_s == ref, "md5": hashlib.md5(ref).hexdigest()}}
            row["speedup_vs_reference"] = round(row["e2e_reads_per_s"] / row["reference"]["reads_per_s"], 1)
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"rows": rows}, open(os.path.join(ROOT, "gpurun_out", "config_bench.json"), "w"), indent=1)
    return 0 if all(r["reference"]["identical_on_sample"] for r in rows) else 1


if __name__ == "__main__":
    sys.exit(main())
