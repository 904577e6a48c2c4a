#!/usr/bin/env python
"""Makes tests/golden/reader_golden.json: tricky FASTA/FASTQ inputs and what the REFERENCE's own reader
(oracle/_ref/kseq_dump = /root/reference/src/kseq.h behind a dump loop, built by oracle/Makefile) reads from them.
Run once in the build container (needs /root/reference); the fixture travels, the reference does not."""
import base64
import gzip
import json
import os
import random
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DUMP = os.path.join(ROOT, "oracle", "_ref", "kseq_dump")


def cases():
    rnd = random.Random(7)

    def dna(n):
        return "".join(rnd.choice("ACGT") for _ in range(n))
    c = {}
    c["fasta_single_line"] = ">r1 comment here\n%s\n>r2\n%s\n" % (dna(50), dna(70))
    c["fasta_multi_line"] = ">r1\n%s\n%s\n%s\n>r2\tx\n%s\n%s\n" % (dna(60), dna(60), dna(13), dna(60), dna(1))
    c["fasta_no_trailing_newline"] = ">r1\n%s\n>r2\n%s" % (dna(40), dna(33))
    c["fasta_crlf"] = ">r1 c\r\n%s\r\n%s\r\n>r2\r\n%s\r\n" % (dna(30), dna(30), dna(5))
    c["fasta_blank_lines_and_junk_before"] = "junk line\n\n>r1\n\n%s\n\n\n%s\n>r2\n\n" % (dna(20), dna(20))
    c["fasta_empty_seq_and_name_only_at_eof"] = ">e1\n>e2\n%s\n>last" % dna(10)
    c["fasta_spaces_and_lowercase"] = ">r1\nacgt ACGT\tnnNN\n A C\n>r2 \n%s\n" % dna(9)
    c["fasta_lone_cr_lines"] = ">r1\n\r\n%s\r\n\r\n>r2\nA\r\n\r" % dna(12)
    c["fasta_gt_inside_line"] = ">r1\nAC>GT@AC+GT\n>r2\nACGT\n"
    q = "".join(chr(33 + rnd.randrange(0, 60)) for _ in range(50))
    s = dna(50)
    c["fastq_simple"] = "@q1 desc\n%s\n+\n%s\n@q2\n%s\n+q2\n%s\n" % (s, q, s[:20], q[:20])
    c["fastq_multiline_with_at_in_quality"] = "@q1\n%s\n%s\n+\n@%s\n%s\n@q2\n%s\n+\n%s\n" % (s[:25], s[25:], q[1:25], "@" + q[26:], s[:10], "@" * 10)
    c["fastq_truncated_quality_mid_chunk"] = "@q1\n%s\n+\n%s\n@bad\nACGTACGT\n+\nIIII\n@q3\n%s\n+\n%s\n" % (s[:10], q[:10], s[:12], q[:12])
    c["fastq_bad_first_record"] = "@bad\nACGTACGT\n+\nIII\n@q2\n%s\n+\n%s\n" % (s[:12], q[:12])
    c["fastq_no_quality_at_eof"] = "@q1\n%s\n+\n%s\n@q2\nACGT\n+" % (s[:10], q[:10])
    c["fastq_crlf"] = "@q1\r\n%s\r\n+\r\n%s\r\n@q2\r\nAC\r\n+\r\nII\r\n" % (s[:10], q[:10])
    c["mixed_fasta_fastq"] = ">f1\n%s\n@q1\n%s\n+\n%s\n>f2\n%s\n" % (dna(8), s[:9], q[:9], dna(7))
    c["empty_file"] = ""
    c["no_records"] = "just text\nno header\n"
    big = []
    for i in range(4100):       # crosses the reference's 4096-read chunk with a bad record right at a chunk start
        if i == 4096:
            big.append("@bad%d\nACGT\n+\nII\n" % i)
        else:
            big.append("@b%d\n%s\n+\n%s\n" % (i, "ACGTA", "IIIII"))
    c["fastq_bad_record_at_chunk_start"] = "".join(big)
    big = []
    for i in range(4100):       # ... and one bad record in the middle of a chunk (dropped, reading goes on)
        if i == 100:
            big.append("@bad%d\nACGT\n+\nII\n" % i)
        else:
            big.append("@b%d\n%s\n+\n%s\n" % (i, "ACGTA", "IIIII"))
    c["fastq_bad_record_mid_chunk_long"] = "".join(big)
    long_lines = ">long1\n" + "\n".join(dna(80) for _ in range(300)) + "\n>long2\n" + dna(30000) + "\n"
    c["fasta_long_reads_buffer_boundaries"] = long_lines     # the test also builds the reader with a 4 KB buffer: many refills
    return c


def main():
    out = {"how": "oracle/_ref/kseq_dump (reference kseq.h, chunks of 4096) on each input; inputs are base64 of the raw bytes; *_gz cases are the same bytes gzip-compressed on the fly by the test", "cases": {}}
    with tempfile.TemporaryDirectory() as td:
        for name, text in cases().items():
            raw = text.encode("latin-1")
            p = os.path.join(td, "in.fx")
            open(p, "wb").write(raw)
            exp = subprocess.run([DUMP, p], stdout=subprocess.PIPE, check=True).stdout
            pz = os.path.join(td, "in.fx.gz")
            with gzip.open(pz, "wb") as f:
                f.write(raw)
            expz = subprocess.run([DUMP, pz], stdout=subprocess.PIPE, check=True).stdout
            assert exp == expz, name
            out["cases"][name] = {"input_b64": base64.b64encode(gzip.compress(raw, 9)).decode(), "expected_b64": base64.b64encode(gzip.compress(exp, 9)).decode(),
                                  "reads": exp.count(b"\n")}
            print(name, len(raw), "bytes ->", exp.count(b"\n"), "reads")
    with open(os.path.join(ROOT, "tests", "golden", "reader_golden.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
