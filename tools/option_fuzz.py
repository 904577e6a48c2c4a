#!/usr/bin/env python
"""Option-space fuzz: random TideHunter option sets on seeded synthetic reads against the unmodified reference binary
(oracle/_ref/TideHunter).
  python tools/option_fuzz.py [n_trials] [seed]          the oracle's C restatement (CPU, needs no GPU)
  python tools/option_fuzz.py [n_trials] [seed] gpu      the product: host layer over the C ABI on cuda:0
Test infrastructure (it executes oracle/_ref); writes gpurun_out/option_fuzz_<engine>.json."""
import gzip
import hashlib
import json
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    n_trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    engine = sys.argv[3] if len(sys.argv) > 3 else "oracle"
    import oracle_py as O
    if engine == "gpu":
        import tidehunter_b200 as T
    from tidehunter_b200 import synth
    with gzip.open(os.path.join(ROOT, "tests", "golden", "golden.json.gz"), "rt") as f:
        ad = json.load(f)["adapters"]
    five, three = ad["five"], ad["three"]
    bad = []
    with tempfile.TemporaryDirectory() as td:
        p5, p3 = os.path.join(td, "5.fa"), os.path.join(td, "3.fa")
        open(p5, "w").write(">5\n%s\n" % five)
        open(p3, "w").write(">3\n%s\n" % three)
        for it in range(n_trials):
            shape = rnd.choice(["short", "short", "r2c2", "long", "indel", "splint", "single"])
            start = rnd.randrange(0, 10 ** 6)
            if shape == "indel":
                names, seqs = synth.gen_long_indel_reads(rnd.randrange(6, 16), start=start)
            elif shape == "single":
                names, seqs = synth.gen_single_copy(rnd.randrange(8, 24), (five, three), start=start)
            elif shape == "splint":
                names, seqs = synth.gen_reads("r2c2", rnd.randrange(3, 8), start=start, adapters=(five, three), three_rc=rnd.random() < 0.7)
            else:
                names, seqs = synth.gen_reads(shape, {"short": 24, "r2c2": 6, "long": 3}[shape], start=start)
            argv, kw = [], {}

            def opt(flag, key, val):
                argv.extend([flag, str(val)]); kw[key] = val
            if rnd.random() < 0.5:
                opt("-k", "k", rnd.choice([5, 6, 8, 10, 12, 15, 16]))
            if rnd.random() < 0.3:
                opt("-w", "w", rnd.choice([1, 2, 3, 5, 10, 20]))
            if rnd.random() < 0.2:
                argv.append("-H"); kw["hpc"] = 1
            if rnd.random() < 0.3:
                opt("-p", "min_p", rnd.choice([2, 5, 10, 30, 100]))
            if rnd.random() < 0.3:
                opt("-P", "max_p", rnd.choice([300, 1000, 3000, 10000, 50000]))
            if rnd.random() < 0.3:
                opt("-c", "min_copy", rnd.choice([2, 3, 4]))
            if rnd.random() < 0.3:
                opt("-e", "max_div", rnd.choice([0.05, 0.1, 0.2, 0.3, 0.5]))
            if rnd.random() < 0.2:
                opt("-m", "min_len", rnd.choice([5, 30, 100, 500]))
            if rnd.random() < 0.2:
                argv.append("-l"); kw["only_longest"] = 1
            if os.environ.get("TH_FUZZ_LINEAR") or (engine == "gpu" and rnd.random() < 0.1):  # abPOA's linear gap mode (oracle restatement; GPU: wide pass)
                o2 = rnd.choice([0, 24]); argv.extend(["-O", "0,%d" % o2]); kw["gap_open1"] = 0; kw["gap_open2"] = o2
            elif rnd.random() < 0.25 or os.environ.get("TH_FUZZ_AFFINE"):
                o1 = rnd.choice([2, 4, 6]); o2 = 0 if os.environ.get("TH_FUZZ_AFFINE") else rnd.choice([0, 12, 24, 40])
                argv.extend(["-O", "%d,%d" % (o1, o2)]); kw["gap_open1"] = o1; kw["gap_open2"] = o2
            if rnd.random() < (0.6 if os.environ.get("TH_FUZZ_AFFINE") or os.environ.get("TH_FUZZ_LINEAR") else 0.15):
                e1 = rnd.choice([1, 2, 3])
                argv.extend(["-E", "%d,1" % e1]); kw["gap_ext1"] = e1; kw["gap_ext2"] = 1
            if shape in ("splint", "single") or rnd.random() < 0.15:
                argv.extend(["-5", p5, "-3", p3]); kw["five_seq"] = five; kw["three_seq"] = three
                if shape == "single" and rnd.random() < 0.8:
                    argv.extend(["-s", "-F"]); kw["single_copy"] = 1; kw["only_full_length"] = 1
                elif rnd.random() < 0.5:
                    argv.append("-F"); kw["only_full_length"] = 1
                if rnd.random() < 0.3:
                    opt("-a", "ada_match_rat", rnd.choice([0.6, 0.7, 0.9]))
            fmt = rnd.choice([1, 2, 2, 3, 4])
            unit = rnd.random() < 0.2 and fmt <= 2
            if unit:
                argv.append("-u"); kw["only_unit"] = 1
            argv.extend(["-f", str(fmt)]); kw["out_fmt"] = fmt
            path = os.path.join(td, "in.fa")
            O.write_fasta(path, names, seqs)
            r = subprocess.run([O.REF_BIN, "-t", "4"] + argv + [path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
            try:
                if engine == "gpu":
                    th = T.TideHunter(device=0, lanes=2, **kw)
                    out = th.run(names, seqs)
                    th.close()
                else:
                    out = O.run_batch(names, seqs, O.default_para(**kw), threads=4)[0]
            except Exception as e:  # noqa: BLE001
                out = ("EXC %s" % e).encode()
            same = r.returncode == 0 and out == r.stdout
            print("%3d %-6s %-60s ref rc %d, %6d bytes, %s" % (it, shape, " ".join(a if not a.startswith("/") else "<ad>" for a in argv), r.returncode, len(r.stdout),
                                                               "same" if same else "DIFFERENT"), flush=True)
            if not same:
                bad.append({"trial": it, "shape": shape, "start": start, "argv": [a if not a.startswith("/") else "<ad>" for a in argv], "ref_rc": r.returncode,
                            "ref_md5": hashlib.md5(r.stdout).hexdigest(), "ours_md5": hashlib.md5(out).hexdigest(), "ref_err": r.stderr.decode()[-200:]})
    rep = {"engine": engine, "trials": n_trials, "seed": int(sys.argv[2]) if len(sys.argv) > 2 else 1, "different": len(bad), "cases": bad[:10]}
    print(json.dumps(rep, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "option_fuzz_%s.json" % engine), "w"), indent=1)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
