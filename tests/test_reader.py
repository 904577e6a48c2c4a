"""host/th_reader.h (the CLI's FASTA/FASTQ(.gz) batch reader) against the reference's own reader: the golden dumps in
tests/golden/reader_golden.json were made with /root/reference/src/kseq.h behind oracle/kseq_dump.c (see
tests/golden/make_reader_golden.py).  Covers multi-line and CRLF records, blank lines, junk before the first header,
empty sequences, header at EOF, '@' inside quality strings, truncated quality blocks (dropped record, chunk quirks of
src/main.c:173-182, 402), gz input, and buffer refills (4 KB buffer build)."""
import base64
import gzip
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dumps(tmp_path_factory):
    d = tmp_path_factory.mktemp("reader")
    out = {}
    for tag, flags in (("default", []), ("smallbuf", ["-DTHR_BUFSZ=4096"])):
        exe = str(d / ("th_reader_dump_" + tag))
        subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-Wall", "-Werror"] + flags + ["-o", exe, os.path.join(ROOT, "host", "th_reader_dump.c"), "-lz"])
        out[tag] = exe
    return d, out


def test_reader_matches_reference_reader(dumps):
    d, exes = dumps
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "reader_golden.json")))
    assert len(g["cases"]) >= 20
    for name, c in g["cases"].items():
        raw = gzip.decompress(base64.b64decode(c["input_b64"]))
        exp = gzip.decompress(base64.b64decode(c["expected_b64"]))
        plain = str(d / "in.fx"); gz = str(d / "in.fx.gz")
        open(plain, "wb").write(raw)
        with gzip.open(gz, "wb") as f:
            f.write(raw)
        for tag, exe in exes.items():
            for path in (plain, gz):
                for batch in ("1000", "1", "4096", "7"):
                    got = subprocess.run([exe, path, batch], stdout=subprocess.PIPE, check=True).stdout
                    assert got == exp, "%s (%s, %s, batch %s)" % (name, tag, os.path.basename(path), batch)


def test_reader_fuzz_against_reference_reader(tmp_path):
    """1,500 random byte soups and semi-structured FASTA/FASTQ files (markers, CR/LF mixes, truncated qualities) through
    host/th_reader.h with a 64-byte buffer and through the reference's own reader (oracle/_ref/kseq_dump, built in the
    container that has /root/reference).  Skipped where that binary does not exist."""
    import random
    dump = os.path.join(ROOT, "oracle", "_ref", "kseq_dump")
    if not os.path.exists(dump):
        pytest.skip("oracle/_ref/kseq_dump not built (needs /root/reference)")
    exe = str(tmp_path / "th_reader_dump_fz")
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-DTHR_BUFSZ=64", "-o", exe, os.path.join(ROOT, "host", "th_reader_dump.c"), "-lz"])
    rnd = random.Random(20260117)
    frag = [b">", b"@", b"+", b"\n", b"\r\n", b"\r", b" ", b"\t", b"ACGT", b"acgtn", b"N", b"IIII", b"@@", b"++", b">x", b"name", b"\n\n", b"A", b"!~"]
    p = str(tmp_path / "f")
    for it in range(1500):
        if it % 3 == 0:
            parts = []
            for r in range(rnd.randrange(0, 6)):
                s = b"".join(rnd.choice([b"ACGT", b"A", b"acg", b"N"]) for _ in range(rnd.randrange(0, 8)))
                eol = rnd.choice([b"\n", b"\r\n"])
                if rnd.random() < 0.5:
                    parts.append(b">r%d" % r + rnd.choice([b"", b" c", b"\tc"]) + eol + s + eol)
                else:
                    q = b"I" * max(0, len(s) + rnd.choice([0, 0, 0, -1, 1]))
                    if rnd.random() < 0.3 and len(q) > 2:
                        q = q[:len(q) // 2] + eol + q[len(q) // 2:]
                    parts.append(b"@r%d" % r + eol + s + eol + b"+" + eol + q + rnd.choice([eol, b""]))
            data = b"".join(parts)
        else:
            data = b"".join(rnd.choice(frag) for _ in range(rnd.randrange(0, 40)))
        with open(p, "wb") as f:
            f.write(data)
        exp = subprocess.run([dump, p], stdout=subprocess.PIPE, check=True).stdout
        got = subprocess.run([exe, p, str(rnd.choice([1, 2, 1000]))], stdout=subprocess.PIPE, check=True).stdout
        assert got == exp, repr(data)
