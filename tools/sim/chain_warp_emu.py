"""Lock-step emulation (32 lanes as numpy vectors) of chain_dp_kernel's bucketed path in tidehunter_b200/csrc/th_chain.cuh,
statement by statement: the stable scatter with __match_any_sync, the 32-ary search for the window's far end, the batch
logic with ballots / shuffles / prefix maxima.  Checked against the oracle.  Development aid: it exists because GPU time
was scarce when the kernel was written; the kernel itself is tested by tests/test_gpu_parity.py::test_stage_parity.
usage: python tools/sim/chain_warp_emu.py [shape] [reads]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for d in ('', 'oracle', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, d))
import helpers as H  # noqa: E402
import oracle_py as O  # noqa: E402
from tidehunter_b200 import synth  # noqa: E402

S, BK_MAX = 512, 64
NO, REG, SAME, OVL = 0, 1, 2, 3
INT_MIN, INT_MAX = -2 ** 31, 2 ** 31 - 1
LANE = np.arange(32)


def ballot(pred):
    return int(sum(1 << i for i in range(32) if pred[i]))


def ffs(m):
    return (m & -m).bit_length()   # 1-based, 0 for m == 0


def shfl_up(v, d):
    out = v.copy()
    out[d:] = v[:32 - d]
    return out


def con_score_vec(cs, ce, ps, pe, K, active):
    """con_score for the active lanes; returns (cls, score) vectors"""
    cls = np.zeros(32, dtype=np.int64)
    sc = np.zeros(32, dtype=np.int64)
    for l in range(32):
        if not active[l]:
            continue
        cp, pp = ce - cs, int(pe[l] - ps[l])
        if cs <= ps[l] or 5 * cp >= 9 * pp or 5 * pp >= 9 * cp:
            continue
        de, ds, dpd = abs(ce - int(pe[l])), abs(cs - int(ps[l])), abs(cp - pp)
        matched = min(de, K) + min(ds, K)
        v = de + ds
        lg = v.bit_length() - 1 if v else -1
        sc[l] = matched - (dpd * dpd // 2 + int(lg / 2))
        cls[l] = (OVL if matched < 2 * K else SAME) if dpd == 0 else REG
    return cls, sc


def chain_read(en, pr, K, max_p, cap):
    n = len(en)
    en = np.array(en, dtype=np.int64)
    pr = np.array(pr, dtype=np.int64)
    sc = K + np.minimum(K, pr)
    fr = -np.ones(n, dtype=np.int64)
    nbk = min(max_p // S + 2, 1 << 20)
    bucketed = nbk <= BK_MAX and 2 * n <= cap
    assert bucketed, "the emulation covers the bucketed path only"
    s_fill = np.zeros(BK_MAX, dtype=np.int64)
    s_off = np.zeros(BK_MAX + 1, dtype=np.int64)
    bi = np.zeros(cap, dtype=np.int64); be = np.zeros(cap, dtype=np.int64); bp = np.zeros(cap, dtype=np.int64); bs = np.zeros(cap, dtype=np.int64)
    for i in range(n):
        b = min(int(pr[i]) // S, nbk - 1)
        s_fill[b] += 1
        if b >= 1:
            s_fill[b - 1] += 1
    acc = 0
    for b in range(nbk):
        c = int(s_fill[b]); s_off[b] = acc; s_fill[b] = acc; acc += c
    s_off[nbk] = acc
    for base in range(0, 2 * n, 32):
        e = base + LANE
        i = e >> 1
        key = -1 - LANE.copy()
        pv = np.zeros(32, dtype=np.int64); ev = np.zeros(32, dtype=np.int64)
        for l in range(32):
            if e[l] < 2 * n:
                pv[l] = pr[i[l]]; ev[l] = en[i[l]]
                b = min(int(pv[l]) // S, nbk - 1) - 1 + int(e[l] & 1)
                if b >= 0:
                    key[l] = b
        m = np.array([ballot(key == key[l]) for l in range(32)])       # __match_any_sync
        below = (1 << LANE) - 1
        pos = np.zeros(32, dtype=np.int64)
        for l in range(32):
            if key[l] >= 0:
                pos[l] = s_fill[key[l]] + bin(int(m[l]) & int(below[l])).count("1")
        for l in range(32):   # after __syncwarp
            if key[l] >= 0 and (int(m[l]) & int(below[l])) == 0:
                s_fill[key[l]] += bin(int(m[l])).count("1")
        for l in range(32):
            if key[l] >= 0:
                bi[pos[l]] = i[l]; be[pos[l]] = ev[l]; bp[pos[l]] = pv[l]; bs[pos[l]] = K + min(K, int(pv[l]))
    assert all(s_fill[b] == s_off[b + 1] for b in range(nbk))
    s_fill[:] = 0
    b = min(int(pr[0]) // S, nbk - 1)
    s_fill[b] = 1
    if b >= 1:
        s_fill[b - 1] = 1
    evals = 0
    gmax = 2 * K
    n_fast = 0
    for cur in range(1, n):
        ce, cp = int(en[cur]), int(pr[cur]); cs = ce - cp
        init = K + min(K, cp)
        max_score, best_pre = init, -1
        max_h = cp
        T2 = 2 * (min(gmax, 1 << 27) + 2 * K)
        Dc = int(np.sqrt(np.float32(T2))) + 1
        fast = False
        wlo = 0
        if Dc <= S // 2:
            lo, hi = 0, cur
            while hi - lo > 32:
                step = (hi - lo + 31) >> 5
                probe = lo + LANE * step
                v = np.array([en[p] if p < hi else INT_MAX for p in probe])
                m = ballot(v >= cs)
                if m == 0:
                    lo = lo + 31 * step + 1
                else:
                    f = ffs(m) - 1
                    hi = min(hi, lo + f * step)
                    if f:
                        lo = lo + (f - 1) * step + 1
            probe = lo + LANE
            v = np.array([en[p] if p < hi else INT_MAX for p in probe])
            m = ballot(v >= cs)
            wlo = min(hi, lo + ffs(m) - 1) if m else hi
            fast = cur - wlo < max_h
        if fast:
            n_fast += 1
            j = max(cp - Dc, 0) // S
            seg_lo = int(s_off[j])
            stop_idx = -1
            base = seg_lo + int(s_fill[j]) - 1
            while base >= seg_lo:
                p = base - LANE
                valid = p >= seg_lo
                pe = np.where(valid, be[np.maximum(p, 0)], 0); pp = np.where(valid, bp[np.maximum(p, 0)], 1)
                psc = np.where(valid, bs[np.maximum(p, 0)], 0); pq = np.where(valid, bi[np.maximum(p, 0)], 0)
                cstop = (~valid) | (pe < cs)
                dpd = np.abs(cp - pp)
                act = (~cstop) & (dpd * dpd < T2)
                cls, con = con_score_vec(cs, ce, pe - pp, pe, K, act)
                s = np.where(cls != NO, psc + con, INT_MIN)
                cm = ballot(cstop)
                first_c = ffs(cm) - 1 if cm else 32
                if s.max() <= max_score:
                    am = ballot((~cstop) & (cls == OVL))
                    first_a = ffs(am) - 1 if am else 32
                else:
                    inc = s.copy()
                    d = 1
                    while d < 32:
                        inc = np.maximum(inc, shfl_up(inc, d)); d <<= 1
                    excl = shfl_up(inc, 1)
                    excl = np.where(LANE == 0, max_score, np.maximum(max_score, excl))
                    imp = (cls != NO) & (s > excl)
                    stop_after = (imp & ((cls == SAME) | (cls == OVL))) | ((~imp) & (cls == OVL))
                    am = ballot(stop_after & (~cstop))
                    first_a = ffs(am) - 1 if am else 32
                    processed = (LANE < first_c) & (LANE <= first_a)
                    bm = int(np.where(processed, s, INT_MIN).max())
                    if bm > max_score:
                        wm = ballot(processed & imp & (s == bm))
                        max_score = bm; best_pre = int(pq[ffs(wm) - 1])
                if first_a < first_c:
                    stop_idx = int(pq[first_a])
                if first_c < 32 or first_a < 32:
                    break
                base -= 32
            evals += cur - max(wlo, stop_idx)
        else:
            iter_n = 0
            for q in range(cur - 1, -1, -1):   # the plain scan (literal)
                if en[q] < cs:
                    break
                evals += 1
                act = np.zeros(32, dtype=bool); act[0] = True
                cls, con = con_score_vec(cs, ce, np.full(32, en[q] - pr[q]), np.full(32, en[q]), K, act)
                if cls[0] != NO:
                    s = int(sc[q] + con[0])
                    if s > max_score:
                        max_score, best_pre = s, q
                        if cls[0] >= SAME:
                            break
                        iter_n = 0
                        continue
                    elif cls[0] == OVL:
                        break
                iter_n += 1
                if iter_n >= max_h:
                    break
        if max_score > init:
            sc[cur] = max_score; fr[cur] = best_pre
        gmax = max(gmax, max_score)
        b = min(cp // S, nbk - 1)
        if max_score > init:
            bs[s_off[b] + s_fill[b]] = max_score
            if b >= 1:
                bs[s_off[b - 1] + s_fill[b - 1]] = max_score
        assert bi[s_off[b] + s_fill[b]] == cur and (b < 1 or bi[s_off[b - 1] + s_fill[b - 1]] == cur)
        s_fill[b] += 1
        if b >= 1:
            s_fill[b - 1] += 1
    return [int(x) for x in sc], [int(x) for x in fr], evals, n_fast


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "r2c2"
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    para = O.default_para()
    bad = 0
    _, seqs = synth.gen_reads(shape, n_reads)
    for seq in seqs:
        hl = H.hits(seq, para)
        en = [h[0] for h in hl]; pr = [h[1] for h in hl]
        if len(hl) < 2 or any(en[i] == en[i - 1] for i in range(1, len(en))) or 2 * len(hl) > len(seq):
            continue
        sc, fr, evals, n_fast = chain_read(en, pr, para.k, para.max_p, (len(seq) + 63) // 64 * 64)
        ref = H.chain(hl, para)
        ok = ref.score == sc and ref.frm == fr and ref.n_evals == evals
        bad += not ok
        print("hits %d fast cells %d evals %d oracle %d %s" % (len(hl), n_fast, evals, ref.n_evals, "ok" if ok else "MISMATCH"))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
