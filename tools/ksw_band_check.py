"""CPU check of the banded identity alignment's certificate (tools/sim/ksw_band_sim.c, the scalar model of
ksw_warp_global2's banded mode): on random unit-vs-consensus pairs of several error rates, repeat structures and length
differences, (1) the full-matrix model equals the oracle's ksw2 restatement, (2) whenever the certificate passes the banded
identity count equals the full one, and it prints the pass rate and the share of cells computed per band fraction.
usage: python tools/ksw_band_check.py [pairs per case]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle_py as O  # noqa: E402
from tidehunter_b200 import synth  # noqa: E402

so = "/tmp/libkswsim.so"
subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "sim", "ksw_band_sim.c")])
S = C.CDLL(so)
S.ksw_band_sim.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_longlong)]
L = O.lib()


def sim(q, t, BW, B, full=False):
    D = len(q) - len(t)
    sc, idn, cells = C.c_int(), C.c_int(), C.c_longlong()
    ok = S.ksw_band_sim(q.ctypes.data, len(q), t.ctypes.data, len(t), BW, B + max(0, D), B + max(0, -D), int(full), C.byref(sc), C.byref(idn), C.byref(cells))
    return ok, sc.value, idn.value, cells.value


def oracle_iden(q, t):
    n = C.c_int()
    return L.tho_ksw2_global(q.ctypes.data, len(q), t.ctypes.data, len(t), C.byref(n), None)


def main():
    npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(7)
    cases = []
    for err, ulen in ((0.10, 120), (0.10, 1000), (0.15, 1000), (0.15, 2400), (0.20, 1000), (0.20, 4500), (0.30, 700)):
        for _ in range(npairs if ulen < 3000 else max(4, npairs // 6)):
            u = rng.integers(0, 4, ulen, dtype=np.uint8)
            cases.append(("err%.2f_len%d" % (err, ulen), synth._channel(rng, u, err), synth._channel(rng, u, 0.01)))
    for _ in range(npairs):  # low complexity / internal repeats: many co-optimal paths
        m = rng.integers(1, 6)
        u = np.tile(rng.integers(0, 4, m, dtype=np.uint8), 800 // m)
        cases.append(("lowcomplex", synth._channel(rng, u, 0.1), synth._channel(rng, u, 0.02)))
        v = np.tile(rng.integers(0, 4, 97, dtype=np.uint8), 8)
        cases.append(("repeat97", synth._channel(rng, v, 0.15), v.copy()))
        a = rng.integers(0, 4, rng.integers(300, 900), dtype=np.uint8)
        b = rng.integers(0, 4, rng.integers(300, 900), dtype=np.uint8)
        cases.append(("unrelated", a, b))
        w = rng.integers(0, 4, 900, dtype=np.uint8)
        k = rng.integers(50, 300)
        cases.append(("bigindel", np.concatenate([w[:400], w[400 + k:]]), w.copy()))
        cases.append(("bigins", w.copy(), np.concatenate([w[:300], w[300 + k:]])))
    stats = {}
    bad = 0
    for name, q, t in cases:
        q = np.ascontiguousarray(q); t = np.ascontiguousarray(t)
        exp = oracle_iden(q, t)
        ok, sc_full, idn, cells_full = sim(q, t, 256, 0, full=True)
        if not ok or idn != exp:
            bad += 1; print("FULL MODEL MISMATCH", name, idn, exp)
        for BW in (256, 512):
            for frac in (0.05, 0.10, 0.15, 0.19, 0.22, 0.25, 0.30):
                B = int(frac * max(len(q), len(t))) + 16
                ok, sc, idn, cells = sim(q, t, BW, B)
                st = stats.setdefault((name, BW, frac), [0, 0, 0.0])
                st[0] += 1; st[1] += ok; st[2] += cells / cells_full
                if ok and (idn != exp or sc != sc_full):
                    bad += 1; print("CERTIFIED BUT DIFFERENT", name, BW, frac, len(q), len(t), idn, exp, sc, sc_full)
                if sc > sc_full:
                    bad += 1; print("BANDED SCORE ABOVE FULL", name, BW, frac)
    for (name, BW, frac), (n, ok, cf) in sorted(stats.items()):
        print("%-18s BW %3d band %.2f: pass %3d / %3d, cells %.2f of full" % (name, BW, frac, ok, n, cf / n))
    print("bad:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
