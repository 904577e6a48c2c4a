"""The certified band of the packed identity alignments (tidehunter_b200/csrc/th_ksw.cuh: ksw_warp_global2's banded mode)
checked on the CPU: tools/sim/ksw_band_sim.c restates the kernel's recurrence, block geometry, lower-bound boundary values
and certificate in scalar C; whenever the certificate passes, the banded score and identity count must equal the oracle's
full-matrix ksw2 (ksw2/ksw2_extz2_sse.c + ksw2_get_xid, src/ksw2_align.c:62-86).  The GPU routine itself is compared with
the oracle in tests/test_gpu_parity.py::test_ksw_banded_pair_identity."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as O
from tidehunter_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("kswsim") / "libkswsim.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tools", "sim", "ksw_band_sim.c")])
    lib = C.CDLL(so)
    lib.ksw_band_sim.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_longlong)]

    def run(q, t, bw, band, full=False):
        d = len(q) - len(t)
        sc, idn, cells = C.c_int(), C.c_int(), C.c_longlong()
        ok = lib.ksw_band_sim(q.ctypes.data, len(q), t.ctypes.data, len(t), bw, band + max(0, d), band + max(0, -d), int(full),
                              C.byref(sc), C.byref(idn), C.byref(cells))
        return ok, sc.value, idn.value, cells.value
    return run


def _cases(rng):
    out = []
    for err, ulen, n in ((0.10, 300, 6), (0.15, 1000, 6), (0.25, 800, 6), (0.15, 2400, 2)):
        for _ in range(n):
            u = rng.integers(0, 4, ulen, dtype=np.uint8)
            out.append((synth._channel(rng, u, err), synth._channel(rng, u, 0.01)))
    for _ in range(5):
        m = int(rng.integers(1, 6))
        u = np.tile(rng.integers(0, 4, m, dtype=np.uint8), 700 // m)            # low complexity: many co-optimal paths
        out.append((synth._channel(rng, u, 0.1), synth._channel(rng, u, 0.02)))
        v = np.tile(rng.integers(0, 4, 97, dtype=np.uint8), 7)                  # internal repeat: shifted alignments score well
        out.append((synth._channel(rng, v, 0.15), v.copy()))
        out.append((rng.integers(0, 4, int(rng.integers(300, 700)), dtype=np.uint8), rng.integers(0, 4, int(rng.integers(300, 700)), dtype=np.uint8)))
        w = rng.integers(0, 4, 800, dtype=np.uint8)
        k = int(rng.integers(40, 300))
        out.append((np.concatenate([w[:350], w[350 + k:]]), w.copy()))          # one long deletion / insertion
        out.append((w.copy(), np.concatenate([w[:300], w[300 + k:]])))
    return [(np.ascontiguousarray(q), np.ascontiguousarray(t)) for q, t in out]


def test_certified_band_equals_full_matrix(sim):
    L = O.lib()
    rng = np.random.default_rng(5)
    certified = failed = 0
    for q, t in _cases(rng):
        n = C.c_int()
        exp = L.tho_ksw2_global(q.ctypes.data, len(q), t.ctypes.data, len(t), C.byref(n), None)
        ok, sc_full, idn, cells_full = sim(q, t, 256, 0, full=True)
        assert ok and idn == exp, ("the model's full matrix differs from the oracle", len(q), len(t), idn, exp)
        for bw in (128, 256, 512):
            for frac in (0.03, 0.10, 0.19, 0.30):
                band = int(frac * max(len(q), len(t))) + 16
                ok, sc, idn, cells = sim(q, t, bw, band)
                assert sc <= sc_full, "a banded score is a lower bound of the full one"
                assert cells <= cells_full
                if ok:
                    certified += 1
                    assert (sc, idn) == (sc_full, exp), ("certified but different", bw, frac, len(q), len(t))
                else:
                    failed += 1
    assert certified > 100 and failed > 100, (certified, failed)   # both outcomes are exercised
