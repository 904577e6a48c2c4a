"""Shared fixtures.  `-m "not gpu"` runs on the CPU-only container; `-m gpu` needs a B200."""
import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "golden.json.gz"), "rt") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_inputs(golden):
    """name -> (names, seqs) for every input set the golden cases refer to."""
    from tidehunter_b200 import synth
    five, three = golden["adapters"]["five"], golden["adapters"]["three"]
    cache = {}

    def get(tag):
        if tag in cache:
            return cache[tag]
        if tag in golden["inputs"]:
            d = golden["inputs"][tag]
            v = ([x.encode() for x in d["names"]], [x.encode() for x in d["seqs"]])
        elif tag == "testfq30":
            n, s = get("testfq_all")
            v = (n[:30], s[:30])
        elif tag == "syn_r2c2":
            v = synth.gen_reads("r2c2", 20)
        elif tag == "syn_short":
            v = synth.gen_reads("short", 20)
        elif tag == "syn_long":
            v = synth.gen_reads("long", 8)
        elif tag == "syn_adapter":
            v = synth.gen_reads("r2c2", 12, adapters=(five, three))
        elif tag == "syn_single":
            v = synth.gen_single_copy(24, (five, three))
        else:
            raise KeyError(tag)
        cache[tag] = v
        return v
    return get


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    oracle_py.lib()
    return oracle_py
