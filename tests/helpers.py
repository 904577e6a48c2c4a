"""Stage-level access to the CPU oracle for the parity tests (ctypes over oracle/libth_oracle.so)."""
import ctypes as C

import numpy as np

import oracle_py as O


def bseq_of(seq):
    L = O.lib()
    buf = (C.c_uint8 * max(len(seq), 1))()
    L.tho_get_bseq(seq, len(seq), buf)
    return bytes(buf[:len(seq)])


def hits(seq, para):
    """[(end, period)] as collect_tandem_repeat_hit would return them."""
    L = O.lib()
    b = bseq_of(seq)
    hp = C.POINTER(C.c_uint64)()
    n = L.tho_collect_hits(b, len(b), C.byref(para), C.byref(hp))
    out = [(int(hp[i] >> 32), int(hp[i] & 0xffffffff)) for i in range(n)]
    if n:
        L.tho_free(hp)
    return out


class ChainResult:
    pass


def chain(hit_list, para):
    L = O.lib()
    n = len(hit_list)
    arr = (C.c_uint64 * max(n, 1))(*[(e << 32) | p for e, p in hit_list])
    ch = O.Chain()
    nch = L.tho_tandem_chain(arr, n, C.byref(para), C.byref(ch))
    r = ChainResult()
    r.raw = ch
    r.n_cells = ch.n_cells
    r.score = [ch.score[i] for i in range(ch.n_cells)]
    r.frm = [ch.from_[i] for i in range(ch.n_cells)]
    r.chains = [[ch.cells[j] for j in range(ch.chain_off[i], ch.chain_off[i + 1])] for i in range(nch)]
    r.n_evals = ch.n_evals
    return r


def partition(seq, chain_res, ci, para):
    L = O.lib()
    b = bseq_of(seq)
    pp = C.POINTER(C.c_int)()
    n = L.tho_partition(b, len(b), C.byref(chain_res.raw), ci, C.byref(para), C.byref(pp))
    out = [pp[i] for i in range(n)]
    L.tho_free(pp)
    return out


def ksw_global(q, t):
    """(iden_n, cigar list) for nt4-coded bytes q (query) and t (target)."""
    L = O.lib()
    nc = C.c_int(0)
    cg = C.POINTER(C.c_uint32)()
    iden = L.tho_ksw2_global(q, len(q), t, len(t), C.byref(nc), C.byref(cg))
    cig = [cg[i] for i in range(nc.value)]
    if nc.value:
        L.tho_free(cg)
    return iden, cig


def ksw_left_end(cig, ql, tl, q_left_ext):
    L = O.lib()
    arr = (C.c_uint32 * max(len(cig), 1))(*cig)
    return L.tho_ksw2_backtrack_left_end(len(cig), arr, ql, tl, q_left_ext)


def ksw_ext(q, t):
    L = O.lib()
    mq, mt = C.c_int(0), C.c_int(0)
    L.tho_ksw2_ext(q, len(q), t, len(t), C.byref(mq), C.byref(mt))
    return mq.value, mt.value


def random_pairs(rng, n, lo, hi, div=0.15, with_n=False):
    """n (query, target) nt4 pairs: target = mutated copy of query."""
    out = []
    for _ in range(n):
        l = int(rng.integers(lo, hi + 1))
        q = rng.integers(0, 4, l, dtype=np.uint8)
        u = rng.random(l)
        keep = u >= div / 3
        t = q.copy()
        sub = (u >= div / 3) & (u < 2 * div / 3)
        t[sub] = (t[sub] + rng.integers(1, 4, int(sub.sum()), dtype=np.uint8)) & 3
        t = t[keep]
        ins = np.flatnonzero(rng.random(len(t)) < div / 3)
        t = np.insert(t, ins, rng.integers(0, 4, len(ins), dtype=np.uint8))
        if with_n:
            q[rng.random(len(q)) < 0.02] = 4
            t[rng.random(len(t)) < 0.02] = 4
        if len(t) == 0:
            t = np.array([0], dtype=np.uint8)
        out.append((q.tobytes(), t.tobytes()))
    return out
