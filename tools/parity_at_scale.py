#!/usr/bin/env python
"""Parity at scale on the GPU box: the unmodified reference binary (oracle/_ref/TideHunter, all host cores) against the
GPU path (host layer over the C ABI) on thousands of seeded synthetic reads of every BASELINE.json shape, byte for byte.

  python tools/parity_at_scale.py [--scale 1.0] [--out gpurun_out/parity_at_scale.json]

One JSON record per case: reads, bases, md5 of both outputs, identical or the first differing read.  Test infrastructure
(it executes oracle/_ref); nothing here is on the product path.
"""
import argparse
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_at_scale.json"))
    args = ap.parse_args()
    import gzip
    import oracle_py as O
    import tidehunter_b200 as T
    from tidehunter_b200 import synth
    with gzip.open(os.path.join(ROOT, "tests", "golden", "golden.json.gz"), "rt") as f:
        ad = json.load(f)["adapters"]
    five, three = ad["five"], ad["three"]
    cores = os.cpu_count() or 1
    sc = args.scale

    def n_of(x):
        return max(int(x * sc), 8)

    cases = [
        ("r2c2 -f 1", lambda: synth.gen_reads("r2c2", n_of(16384), start=200000), ["-f", "1"], dict(out_fmt=1)),
        ("r2c2 -f 4", lambda: synth.gen_reads("r2c2", n_of(6000), start=300000), ["-f", "4"], dict(out_fmt=4)),
        ("short -f 2", lambda: synth.gen_reads("short", n_of(16384), start=200000), ["-f", "2"], dict(out_fmt=2)),
        ("long -f 2", lambda: synth.gen_reads("long", n_of(3072), start=200000), ["-f", "2"], dict(out_fmt=2)),
        ("r2c2 -u -f 2", lambda: synth.gen_reads("r2c2", n_of(8192), start=400000), ["-u", "-f", "2"], dict(out_fmt=2, only_unit=1)),
        ("adapter -5 -3 -f 2", lambda: synth.gen_reads("r2c2", n_of(4096), start=200000, adapters=(five, three)), ["ADAPTERS", "-f", "2"],
         dict(out_fmt=2, five_seq=five, three_seq=three)),
        ("adapter -5 -3 -F -f 2", lambda: synth.gen_reads("r2c2", n_of(2048), start=210000, adapters=(five, three), three_rc=True), ["ADAPTERS", "-F", "-f", "2"],
         dict(out_fmt=2, five_seq=five, three_seq=three, only_full_length=1)),
        ("splint -5 -3 -f 3", lambda: synth.gen_reads("r2c2", n_of(3000), start=220000, adapters=(five, three), three_rc=True), ["ADAPTERS", "-f", "3"],
         dict(out_fmt=3, five_seq=five, three_seq=three)),
        ("single -s -F -f 2", lambda: synth.gen_single_copy(n_of(4096), (five, three), start=200000), ["ADAPTERS", "-s", "-F", "-f", "2"],
         dict(out_fmt=2, five_seq=five, three_seq=three, only_full_length=1, single_copy=1)),
        ("r2c2 -k 12 -w 5 -f 2", lambda: synth.gen_reads("r2c2", n_of(4096), start=500000), ["-k", "12", "-w", "5", "-f", "2"], dict(out_fmt=2, k=12, w=5)),
        ("short -p 10 -c 3 -f 2", lambda: synth.gen_reads("short", n_of(4096), start=500000), ["-p", "10", "-c", "3", "-f", "2"], dict(out_fmt=2, min_p=10, min_copy=3)),
    ]
    report = {"cores": cores, "cases": []}
    ok = True
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        p5, p3 = os.path.join(td, "5.fa"), os.path.join(td, "3.fa")
        open(p5, "w").write(">5\n%s\n" % five)
        open(p3, "w").write(">3\n%s\n" % three)
        for tag, gen, argv, kw in cases:
            names, seqs = gen()
            path = os.path.join(td, "in.fa")
            O.write_fasta(path, names, seqs)
            argv = sum((["-5", p5, "-3", p3] if a == "ADAPTERS" else [a] for a in argv), [])
            t0 = time.perf_counter()
            ref = subprocess.run([O.REF_BIN, "-t", str(cores)] + argv + [path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            t_ref = time.perf_counter() - t0
            th = T.TideHunter(device=0, **kw)
            t0 = time.perf_counter()
            ours = th.run(names, seqs)
            t_gpu = time.perf_counter() - t0
            th.close()
            rec = {"case": tag, "reads": len(seqs), "bases": synth.total_bases(seqs), "identical": ours == ref, "bytes": len(ref),
                   "md5_reference": hashlib.md5(ref).hexdigest(), "md5_ours": hashlib.md5(ours).hexdigest(),
                   "reference_s": round(t_ref, 2), "ours_s_incl_first_call": round(t_gpu, 2)}
            if ours != ref:
                ok = False
                a, b = ref.split(b"\n"), ours.split(b"\n")
                i = next((i for i in range(min(len(a), len(b))) if a[i] != b[i]), min(len(a), len(b)))
                rec["first_diff_line"] = i
                rec["reference_line"] = a[i][:200].decode(errors="replace") if i < len(a) else None
                rec["ours_line"] = b[i][:200].decode(errors="replace") if i < len(b) else None
            report["cases"].append(rec)
            print(json.dumps(rec), flush=True)
    report["all_identical"] = ok
    report["total_reads"] = sum(c["reads"] for c in report["cases"])
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)
    print("ALL IDENTICAL" if ok else "MISMATCH", report["total_reads"], "reads")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
