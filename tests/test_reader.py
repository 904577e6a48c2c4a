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
