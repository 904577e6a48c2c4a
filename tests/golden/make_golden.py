"""Generates tests/golden/golden.json.gz by running the UNMODIFIED reference (oracle/_ref/TideHunter,
built by oracle/Makefile with -fno-strict-aliasing -mavx2) in this container.

Inputs: the reference's own smoke inputs (test_data/*.fa, the first 30 reads of test.fq) and seeded
synthetic reads of the three BASELINE shapes (tidehunter_b200/synth.py).  The input reads of the
reference's files are stored in the fixture too, because /root/reference does not exist on the GPU box.

Run:  python tests/golden/make_golden.py
"""
import gzip
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
from tidehunter_b200 import synth  # noqa: E402

REF = "/root/reference"


def main():
    O.build()
    sets = {}
    for tag, fn in (("test_50x4", "test_data/test_50x4.fa"), ("test_1000x10", "test_data/test_1000x10.fa"), ("full_length", "test_data/full_length.fa")):
        n, s = O.read_fastx(os.path.join(REF, fn))
        sets[tag] = (n, s)
    n, s = O.read_fastx(os.path.join(REF, "test.fq"))
    sets["testfq30"] = (n[:30], s[:30])
    sets["testfq_all"] = (n, s)  # all 100 reads are stored; testfq30 = their first 30 (derived by the tests)
    sets["syn_r2c2"] = synth.gen_reads("r2c2", 20)
    sets["syn_short"] = synth.gen_reads("short", 20)
    sets["syn_long"] = synth.gen_reads("long", 8)
    five = O.read_fastx(os.path.join(REF, "test_data/5prime.fa"))[1][0].decode()
    three = O.read_fastx(os.path.join(REF, "test_data/3prime.fa"))[1][0].decode()
    sets["syn_adapter"] = synth.gen_reads("r2c2", 12, adapters=(five, three))

    cases = []  # (input set, reference argv, oracle/para keywords)
    for st in ("test_50x4", "test_1000x10", "testfq30", "syn_r2c2", "syn_short", "syn_long"):
        for fmt in (1, 2, 3, 4):
            if st.startswith("syn") and fmt in (1, 3):
                continue
            cases.append((st, ["-f", str(fmt)], dict(out_fmt=fmt)))
    for st in ("testfq30", "syn_r2c2"):
        cases.append((st, ["-u", "-f", "1"], dict(out_fmt=1, only_unit=1)))
        cases.append((st, ["-u", "-f", "2"], dict(out_fmt=2, only_unit=1)))
    for args, kw in ((["-w", "5"], dict(w=5)), (["-w", "10", "-k", "12"], dict(w=10, k=12)), (["-k", "15"], dict(k=15)),
                     (["-k", "16"], dict(k=16)), (["-H"], dict(hpc=1)), (["-p", "2"], dict(min_p=2)), (["-H", "-w", "4"], dict(hpc=1, w=4)),
                     (["-c", "3", "-e", "0.1"], dict(min_copy=3, max_div=0.1)), (["-l"], dict(only_longest=1)),
                     (["-m", "500", "-P", "2000"], dict(min_len=500, max_p=2000))):
        cases.append(("testfq30", args + ["-f", "2"], dict(out_fmt=2, **kw)))
    for st in ("full_length", "syn_adapter"):
        ad = ["-5", os.path.join(REF, "test_data/5prime.fa"), "-3", os.path.join(REF, "test_data/3prime.fa")]
        cases.append((st, ad + ["-f", "2"], dict(out_fmt=2, five_seq=five, three_seq=three)))
        cases.append((st, ad + ["-F", "-f", "2"], dict(out_fmt=2, five_seq=five, three_seq=three, only_full_length=1)))
        cases.append((st, ad + ["-u", "-f", "2"], dict(out_fmt=2, five_seq=five, three_seq=three, only_unit=1)))
    for fmt in (1, 2, 3, 4):
        cases.append(("testfq_all", ["-f", str(fmt)], dict(out_fmt=fmt)))
    # -s: single-copy full-length reads (src/gen_cons.c:128-171), alone and next to tandem records
    sets["syn_single"] = synth.gen_single_copy(24, (five, three))
    ad = ["-5", os.path.join(REF, "test_data/5prime.fa"), "-3", os.path.join(REF, "test_data/3prime.fa")]
    akw = dict(five_seq=five, three_seq=three, only_full_length=1, single_copy=1)
    for st in ("syn_single", "full_length"):
        cases.append((st, ad + ["-s", "-F", "-f", "2"], dict(out_fmt=2, **akw)))
    cases.append(("syn_adapter", ad + ["-s", "-F", "-a", "0.6", "-f", "2"], dict(out_fmt=2, ada_match_rat=0.6, **akw)))
    cases.append(("syn_single", ad + ["-s", "-F", "-f", "1"], dict(out_fmt=1, **akw)))
    cases.append(("syn_single", ad + ["-s", "-F", "-f", "4"], dict(out_fmt=4, **akw)))
    cases.append(("syn_single", ad + ["-s", "-F", "-u", "-f", "1"], dict(out_fmt=1, only_unit=1, **akw)))
    cases.append(("syn_single", ad + ["-s", "-F", "-l", "-f", "2"], dict(out_fmt=2, only_longest=1, **akw)))
    cases.append(("syn_single", ad + ["-s", "-F", "-a", "0.9", "-m", "1000", "-f", "2"], dict(out_fmt=2, ada_match_rat=0.9, min_len=1000, **akw)))

    out = {"inputs": {}, "cases": [], "adapters": {"five": five, "three": three}}
    for tag, (n, s) in sets.items():
        if tag.startswith("syn") or tag == "testfq30":
            continue
        out["inputs"][tag] = {"names": [x.decode() for x in n], "seqs": [x.decode() for x in s]}
    with tempfile.TemporaryDirectory() as td:
        for st, args, kw in cases:
            n, s = sets[st]
            path = os.path.join(td, st + ".fa")
            O.write_fasta(path, n, s)
            ref = O.run_ref(path, args, threads=8)
            printable = [a if not a.startswith(REF) else os.path.basename(a) for a in args]
            rec = {"input": st, "args": printable, "para": kw, "md5": hashlib.md5(ref).hexdigest(), "n_reads": len(n)}
            if st != "testfq_all":
                rec["text"] = ref.decode()
            out["cases"].append(rec)
            print(st, printable, rec["md5"], len(ref))
    with gzip.open(os.path.join(ROOT, "tests", "golden", "golden.json.gz"), "wt") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
