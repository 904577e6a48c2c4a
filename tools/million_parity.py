#!/usr/bin/env python
"""Bit-exact check on the FULL BASELINE.json configs[1] workload: 1,048,576 synthetic R2C2 reads.

The unmodified reference needs ~1 CPU-hour for this set, so the work is split in two:

  python tools/million_parity.py --make-md5      # no GPU: runs oracle/_ref/TideHunter over the 64 chunks of 16,384
                                                 # reads and writes tests/golden/r2c2_1m_md5.json (md5 + bytes per chunk)
  python tools/million_parity.py --check         # GPU box: regenerates the same seeded reads, runs them through the host
                                                 # layer over the C ABI and compares every chunk's md5 with the fixture

Reads are a pure function of (seed, shape, index) (tidehunter_b200/synth.py), default options, -f 1 (FASTA; independent
of the chunking).  Test infrastructure: --make-md5 executes oracle/_ref; nothing here is on the product path.
"""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FIXTURE = os.path.join(ROOT, "tests", "golden", "r2c2_1m_md5.json")
CHUNK = 16384
N_CHUNKS = 64


def _gen_part(a):
    from tidehunter_b200 import synth
    return synth.gen_reads("r2c2", a[1], start=a[0])


def gen_chunk(pool, ci, parts=16):
    step = CHUNK // parts
    res = pool.map(_gen_part, [(ci * CHUNK + i * step, step) for i in range(parts)])
    names, seqs = [], []
    for n, q in res:
        names += n
        seqs += q
    return names, seqs


def make_md5(args):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    O.build()
    done = {}
    if os.path.exists(FIXTURE):
        done = {c["chunk"]: c for c in json.load(open(FIXTURE))["chunks"]}
    cores = os.cpu_count() or 1
    with mp.Pool(min(cores, 8)) as pool, tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        for ci in range(args.first, args.last):
            if ci in done:
                continue
            names, seqs = gen_chunk(pool, ci)
            path = os.path.join(td, "in.fa")
            O.write_fasta(path, names, seqs)
            t0 = time.perf_counter()
            ref = subprocess.run([O.REF_BIN, "-t", str(cores), "-f", "1", path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            done[ci] = {"chunk": ci, "first_read": ci * CHUNK, "reads": len(seqs), "bases": sum(map(len, seqs)), "records": ref.count(b">"),
                        "bytes": len(ref), "md5": hashlib.md5(ref).hexdigest(), "input_md5": hashlib.md5(b"".join(seqs)).hexdigest()}
            print(json.dumps(done[ci]), "%.0f s" % (time.perf_counter() - t0), flush=True)
            with open(FIXTURE + ".tmp", "w") as f:
                json.dump({"workload": "BASELINE.json configs[1]: synth.gen_reads('r2c2', ...) indices 0..1048575, seed %d" % 20260117,
                           "command": "oracle/_ref/TideHunter -f 1 (defaults), unmodified reference built by oracle/Makefile ref",
                           "chunk_reads": CHUNK, "chunks": [done[k] for k in sorted(done)]}, f, indent=0)
            os.replace(FIXTURE + ".tmp", FIXTURE)
    return 0


def check(args):
    import tidehunter_b200 as T
    fx = json.load(open(FIXTURE))
    chunks = [c for c in fx["chunks"] if args.first <= c["chunk"] < args.last]
    th = T.TideHunter(device=0, out_fmt=1)
    ok, reads, bases, t_gpu = True, 0, 0, 0.0
    bad = []
    t_all = time.perf_counter()
    with mp.Pool(min(os.cpu_count() or 1, 32)) as pool:
        nxt = pool.map_async(_gen_part, [(chunks[0]["first_read"] + i * 1024, 1024) for i in range(CHUNK // 1024)]) if chunks else None
        for k, c in enumerate(chunks):
            res = nxt.get()
            if k + 1 < len(chunks):
                nxt = pool.map_async(_gen_part, [(chunks[k + 1]["first_read"] + i * 1024, 1024) for i in range(CHUNK // 1024)])
            names, seqs = [], []
            for n, q in res:
                names += n
                seqs += q
            if hashlib.md5(b"".join(seqs)).hexdigest() != c["input_md5"]:
                raise SystemExit("chunk %d: regenerated reads differ from the fixture's (generator drift)" % c["chunk"])
            t0 = time.perf_counter()
            out = th.run(names, seqs)
            t_gpu += time.perf_counter() - t0
            same = hashlib.md5(out).hexdigest() == c["md5"] and len(out) == c["bytes"]
            if not same:
                ok = False
                bad.append(c["chunk"])
            reads += len(seqs)
            bases += c["bases"]
            print("chunk %2d reads %d..%d %s" % (c["chunk"], c["first_read"], c["first_read"] + len(seqs) - 1, "identical" if same else "MISMATCH"), flush=True)
    th.close()
    rep = {"workload": fx["workload"], "reads": reads, "bases": bases, "chunks": len(chunks), "all_identical": ok, "mismatching_chunks": bad,
           "reference_records": sum(c["records"] for c in chunks), "reference_bytes": sum(c["bytes"] for c in chunks),
           "gpu_path_s": round(t_gpu, 2), "gpu_path_reads_per_s": round(reads / t_gpu, 1) if t_gpu else None,
           "wall_s_incl_read_generation": round(time.perf_counter() - t_all, 1)}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rep, open(args.out, "w"), indent=1)
    print(json.dumps(rep))
    return 0 if ok else 1


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--make-md5", action="store_true")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--last", type=int, default=N_CHUNKS)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "million_parity.json"))
    a = ap.parse_args()
    sys.exit(make_md5(a) if a.make_md5 else check(a))
