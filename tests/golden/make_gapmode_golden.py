#!/usr/bin/env python
"""Makes tests/golden/gapmode_golden.json: outputs of the unmodified reference (oracle/_ref/TideHunter) on seeded reads
with long indels (synth.gen_long_indel_reads) in abPOA's convex gap mode (default -O 4,24 -E 2,1) and in its affine gap
mode (-O 4,0).  The sample separates the two modes.  Run once in the build container (needs /root/reference)."""
import base64
import gzip
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
    n = 160
    names, seqs = synth.gen_long_indel_reads(n)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "in.fa")
        O.write_fasta(p, names, seqs)
        for tag, args in (("convex", []), ("affine", ["-O", "4,0"]), ("affine_e", ["-O", "6,0", "-E", "3,0"])):
            out[tag] = subprocess.run([O.REF_BIN, "-t", "8", "-f", "2"] + args + [p], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    a, b = out["convex"].split(b"\n"), out["affine"].split(b"\n")
    ndiff = sum(1 for i in range(min(len(a), len(b))) if a[i] != b[i])
    assert ndiff > 10, ndiff
    fx = {"how": "oracle/_ref/TideHunter -f 2 [mode args] on synth.gen_long_indel_reads(%d)" % n, "n_reads": n, "lines_differing_convex_vs_affine": ndiff,
          "modes": {"convex": {"args": [], "para": {}}, "affine": {"args": ["-O", "4,0"], "para": {"gap_open2": 0}},
                    "affine_e": {"args": ["-O", "6,0", "-E", "3,0"], "para": {"gap_open1": 6, "gap_open2": 0, "gap_ext1": 3, "gap_ext2": 0}}}}
    for tag in out:
        fx["modes"][tag]["md5"] = hashlib.md5(out[tag]).hexdigest()
        if tag == "affine":     # the text of one mode is kept for diagnostics; the others are pinned by md5
            fx["modes"][tag]["text_gz_b64"] = base64.b64encode(gzip.compress(out[tag], 9)).decode()
    json.dump(fx, open(os.path.join(ROOT, "tests", "golden", "gapmode_golden.json"), "w"), indent=0)
    print(ndiff, {k: v["md5"] for k, v in fx["modes"].items()})


if __name__ == "__main__":
    main()
