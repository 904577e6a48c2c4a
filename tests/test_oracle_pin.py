"""Pins the CPU oracle (oracle/th_*.c, the restatement of the reference's hot path) against

  (a) the reference's own known answers: README.md:222 (test_50x4) and the stdout md5s of SURVEY.md
      section 8(c) for test.fq -f 1..4, test_data/test_1000x10.fa and test_data/test_50x4.fa, and
  (b) outputs of the UNMODIFIED reference compiled here into oracle/_ref/ (tests/golden/golden.json.gz,
      made by tests/golden/make_golden.py): 42 runs over the reference's smoke inputs and seeded
      synthetic reads of every BASELINE.json shape, covering -f 1..4, -u, -k/-w/-H/-p/-c/-e/-l/-m/-P,
      -5/-3/-F.

CPU only; bit-exact (the output text must be byte-identical).
"""
import hashlib

import pytest

# stdout md5s of the deterministic reference build
SURVEY_MD5 = {
    ("testfq_all", 1): "ea70638718813467562e23a51f18f25c",
    ("testfq_all", 2): "9e6da9cab0f872b9899ae238441de647",
    ("testfq_all", 3): "341e96d9cce710af5390a031bd04a92a",
    ("testfq_all", 4): "988ae3d5f90da293dd7acea56fc84513",
    ("test_1000x10", 1): "6518be7cff7c168de0e4c42465b3a6dd",
    ("test_50x4", 1): "66d467e9b90a6e84560354200a9f5233",
}
# /root/reference/README.md:222
README_50x4 = ("test_50x4\trep0\t4.0\t300\t51\t250\t50\t100.0\t0\t59,109,159,208\t"
               "CGATCGATCGGCATGCATGCATGCTAGTCGATGCATCGGGATCAGCTAGT\n")
# SURVEY.md section 8(c), tabular line of BASELINE config 1
CONFIG1_TAB_PREFIX = "test_1000x10\trep0\t9.6\t9710\t101\t9710\t1000\t88.7\t0\t163,1159,2166,3148,4163,5160,6170,7158,8166,9175\t"


def _run(oracle, inputs, para, threads=8):
    names, seqs = inputs
    return oracle.run_batch(names, seqs, oracle.default_para(**para), threads=threads)[0]


def test_readme_known_answer(oracle, golden_inputs):
    assert _run(oracle, golden_inputs("test_50x4"), dict(out_fmt=2)).decode() == README_50x4


def test_config1_tab_line(oracle, golden_inputs):
    assert _run(oracle, golden_inputs("test_1000x10"), dict(out_fmt=2)).decode().startswith(CONFIG1_TAB_PREFIX)


@pytest.mark.parametrize("key", sorted(SURVEY_MD5))
def test_survey_md5(oracle, golden_inputs, key):
    tag, fmt = key
    out = _run(oracle, golden_inputs(tag), dict(out_fmt=fmt))
    assert hashlib.md5(out).hexdigest() == SURVEY_MD5[key]


def test_golden_fixture_agrees_with_survey(golden):
    for c in golden["cases"]:
        key = (c["input"], c["para"].get("out_fmt"))
        if key in SURVEY_MD5 and set(c["para"]) == {"out_fmt"}:
            assert c["md5"] == SURVEY_MD5[key]


def test_all_golden_cases(oracle, golden, golden_inputs):
    bad = []
    for c in golden["cases"]:
        out = _run(oracle, golden_inputs(c["input"]), c["para"])
        if hashlib.md5(out).hexdigest() != c["md5"]:
            bad.append((c["input"], c["args"]))
        elif "text" in c:
            assert out.decode() == c["text"]
    assert not bad, bad


def test_thread_count_does_not_change_output(oracle, golden_inputs):
    inp = golden_inputs("testfq30")
    assert _run(oracle, inp, dict(out_fmt=2), threads=1) == _run(oracle, inp, dict(out_fmt=2), threads=5)


def test_edge_cases(oracle):
    # empty batch, empty read, read shorter than k, all-N read, no repeat at all
    assert oracle.run_batch([], [], oracle.default_para())[0] == b""
    names = [b"e", b"s", b"n", b"u"]
    seqs = [b"", b"ACGT", b"N" * 500, b"ACGTTGCA" * 4 + b"GATTACAGATTACCA"]
    assert oracle.run_batch(names, seqs, oracle.default_para())[0] == b""


def test_million_read_fixture_is_complete():
    """tests/golden/r2c2_1m_md5.json covers BASELINE configs[1] end to end: 64 consecutive chunks of 16,384 reads."""
    import json
    import os
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "r2c2_1m_md5.json")))
    ch = fx["chunks"]
    assert [c["chunk"] for c in ch] == list(range(64))
    assert all(c["first_read"] == 16384 * c["chunk"] and c["reads"] == 16384 for c in ch)
    assert sum(c["reads"] for c in ch) == 1048576 and sum(c["records"] for c in ch) > 1000000
    assert len({c["md5"] for c in ch}) == 64


def _pn8_sample():
    import json
    import os
    from tidehunter_b200 import synth
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pn8_golden.json")))
    names, seqs = [], []
    for shape, start, n in fx["sets"]:
        a, b = synth.gen_reads(shape, n, start=start)
        names += a; seqs += b
    return fx, names, seqs


def test_oracle_sse_vector_width_matches_sse_build_of_the_reference(oracle):
    """pn16 = 8 (abPOA compiled for SSE4.1: 8 int16 lanes per vector) changes band rounding and the row arg-max
    tie-break; the fixture holds the output of the reference built that way (tests/golden/make_pn8_golden.py) on a
    sample that separates the two widths."""
    import hashlib
    fx, names, seqs = _pn8_sample()
    out8 = oracle.run_batch(names, seqs, oracle.default_para(out_fmt=2, pn16=8), threads=4)[0]
    out16 = oracle.run_batch(names, seqs, oracle.default_para(out_fmt=2), threads=4)[0]
    assert out8.decode() == fx["text_pn8"] and hashlib.md5(out8).hexdigest() == fx["md5_pn8"]
    assert hashlib.md5(out16).hexdigest() == fx["md5_pn16"] and out8 != out16


def _gapmode():
    import json
    import os
    from tidehunter_b200 import synth
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gapmode_golden.json")))
    names, seqs = synth.gen_long_indel_reads(fx["n_reads"])
    return fx, names, seqs


def test_oracle_gap_modes_match_reference(oracle):
    """abPOA's convex (default) and affine (-O x,0) gap modes on reads with 22-59 bp indels, where the two modes give
    different consensus sequences (tests/golden/make_gapmode_golden.py ran the unmodified reference)."""
    import hashlib
    fx, names, seqs = _gapmode()
    assert fx["lines_differing_convex_vs_affine"] > 10
    for tag, m in list(fx["modes"].items()) + list(fx["oracle_only_modes"].items()):   # + the linear mode (-O 0,...)
        out = oracle.run_batch(names, seqs, oracle.default_para(out_fmt=2, **m["para"]), threads=4)[0]
        assert hashlib.md5(out).hexdigest() == m["md5"], tag


def test_oracle_int32_path_matches_reference(oracle, capfd, monkeypatch):
    """A 98 kb read whose graph outgrows abPOA's int16 score range: the last alignments run with 32-bit vectors (pn = 8)
    in the reference; the oracle follows and reproduces the reference's record (tests/golden/int32_golden.json)."""
    import hashlib
    import json
    import os
    from tidehunter_b200 import synth
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "int32_golden.json")))
    monkeypatch.setenv("THO_DEBUG", "1")
    names, seqs = synth.gen_int32_read()
    out = oracle.run_batch(names, seqs, oracle.default_para(out_fmt=2), threads=1)[0]
    assert hashlib.md5(out).hexdigest() == fx["md5"] and len(out) == fx["bytes"]
    assert "int32 alignment" in capfd.readouterr().err


def _fuzz_found():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fuzz_found_golden.json")))["cases"]


def test_option_sets_found_by_fuzzing(oracle):
    """Option sets where tools/option_fuzz.py once separated the oracle from the reference: -l replacing a record while a
    quality format is printed (the reference rewinds seq.l but not qual.l, so the replaced record's quality bytes are
    printed for the kept one)."""
    import hashlib
    from tidehunter_b200 import synth
    for c in _fuzz_found():
        names, seqs = synth.gen_reads(c["shape"], c["n"], start=c["start"])
        out = oracle.run_batch(names, seqs, oracle.default_para(**c["para"]), threads=4)[0]
        assert hashlib.md5(out).hexdigest() == c["md5"], c["args"]
