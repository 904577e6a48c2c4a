#!/usr/bin/env python
"""Throughput of the command line front end on a FASTA file in tmpfs (reader thread + host layer + C ABI + fwrite),
next to the reference binary on a prefix of the same file.  usage: python tools/cli_bench.py [n_reads] [ref_reads]"""
import hashlib
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _gen(a):
    from tidehunter_b200 import synth
    return synth.gen_reads("r2c2", a[1], start=a[0])


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    nref = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    from tidehunter_b200 import build as B
    cli = B.build()[2]
    d = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path, pref = os.path.join(d, "th_cli_bench.fa"), os.path.join(d, "th_cli_bench_ref.fa")
    with mp.Pool(min(os.cpu_count() or 1, 32)) as pool, open(path, "wb") as f, open(pref, "wb") as g:
        done = 0
        for names, seqs in pool.imap(_gen, [(i, 1024) for i in range(0, n, 1024)]):
            for nm, s in zip(names, seqs):
                rec = b">" + nm + b"\n" + s + b"\n"
                f.write(rec)
                if done < nref:
                    g.write(rec)
                done += 1
    size = os.path.getsize(path)
    rep = {"reads": n, "fasta_bytes": size}
    devs = os.environ.get("TH_CLI_DEVICES")     # e.g. "0,1": one process drives several GPUs
    extra = ["--devices", devs] if devs else []
    rep["devices"] = devs or "0"
    for tag, cmd in (("cli_fa", [cli] + extra + ["-f", "1", "-o", os.path.join(d, "th_cli_out.fa"), path]),):
        subprocess.run(cmd[:1] + ["-f", "1", "-o", os.devnull, pref], check=True, stderr=subprocess.DEVNULL)   # warm-up (context creation, allocations)
        t0 = time.perf_counter()
        r = subprocess.run(cmd, stderr=subprocess.PIPE, check=True)
        dt = time.perf_counter() - t0
        rep[tag] = {"seconds_incl_process_start": round(dt, 2), "reads_per_s": round(n / dt, 1), "MB_per_s_input": round(size / dt / 1e6, 1), "stderr": r.stderr.decode().strip()[-1500:]}
    # parity of the prefix: CLI vs the reference binary
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "TideHunter")
    if os.path.exists(ref_bin):
        t0 = time.perf_counter()
        ref = subprocess.run([ref_bin, "-t", str(os.cpu_count() or 1), "-f", "1", pref], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        t_ref = time.perf_counter() - t0
        ours = subprocess.run([cli] + extra + ["-f", "1", pref], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        rep["prefix_parity"] = {"reads": nref, "identical": ours == ref, "md5": hashlib.md5(ref).hexdigest(), "reference_reads_per_s": round(nref / t_ref, 1), "cores": os.cpu_count()}
    for p in (path, pref, os.path.join(d, "th_cli_out.fa")):
        try:
            os.unlink(p)
        except OSError:
            pass
    print(json.dumps(rep))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "cli_bench_dev%s.json" % (devs or "0").replace(",", "_")), "w"), indent=1)


if __name__ == "__main__":
    main()
