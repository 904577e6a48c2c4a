#!/usr/bin/env python
"""Makes tests/golden/pn8_golden.json: output of the unmodified reference built with SSE4.1-wide abPOA vectors
(oracle/Makefile ref_sse -> oracle/_ref/sse/TideHunter, pn = 8 int16 lanes) on seeded synthetic reads, including one
read (short #7021) whose consensus differs from the AVX2 build's (pn = 16).  Pins the oracle's pn16 = 8 mode and the
GPU path's simd_lanes16 = 8 mode.  Run once in the build container (needs /root/reference)."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import oracle_py as O
    from tidehunter_b200 import synth
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "ref_sse"])
    sets = [("short", 7000, 48), ("r2c2", 5000, 6), ("long", 5000, 3)]
    names, seqs = [], []
    for shape, start, n in sets:
        a, b = synth.gen_reads(shape, n, start=start)
        names += a; seqs += b
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "in.fa")
        O.write_fasta(p, names, seqs)
        out = {}
        for tag, exe in (("pn8", os.path.join(ROOT, "oracle", "_ref", "sse", "TideHunter")), ("pn16", O.REF_BIN)):
            out[tag] = subprocess.run([exe, "-t", "4", "-f", "2", p], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    assert out["pn8"] != out["pn16"], "the sample no longer separates the two vector widths"
    fx = {"how": "oracle/_ref/sse/TideHunter -f 2 (abPOA SIMD files compiled with -msse4.1 only: pn = 8)", "sets": sets,
          "md5_pn8": hashlib.md5(out["pn8"]).hexdigest(), "md5_pn16": hashlib.md5(out["pn16"]).hexdigest(), "text_pn8": out["pn8"].decode()}
    json.dump(fx, open(os.path.join(ROOT, "tests", "golden", "pn8_golden.json"), "w"), indent=0)
    print(fx["md5_pn8"], fx["md5_pn16"], len(out["pn8"]))


if __name__ == "__main__":
    main()
