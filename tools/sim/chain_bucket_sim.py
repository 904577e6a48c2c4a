"""Model of the period-bucketed chaining DP (th_chain.cuh, chain_dp_kernel's fast path), checked against the oracle.

A predecessor q can only raise the running maximum of cell `cur` when (period(cur) - period(q))^2 < 2 (gmax + 2k), gmax = the
largest score of the read so far (con_score's penalty floor(dpd^2 / 2) would otherwise exceed every score), and the two stop
rules that do not need an improvement (overlap: dpd == 0) cannot fire for it either.  Hits are therefore kept a second time
in period buckets [512 j, 512 j + 1024) (stride 512: every hit is in two of them, in index order), and `cur` only scans the
one bucket that contains [period - D, period + D].  The count of evaluated predecessors the reference would have made
follows from indices (one hit per end position).  Falls back to the plain scan when the window is as long as the period
(the reference's max_h rule could fire), when D > 256, or with more than 64 buckets.
usage: python tools/sim/chain_bucket_sim.py [shape] [reads]"""
import bisect
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for d in ('', 'oracle', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, d))
import helpers as H  # noqa: E402
import oracle_py as O  # noqa: E402
from tidehunter_b200 import synth  # noqa: E402

S = 512
NBK_MAX = 64
NO, REG, SAME, OVL = 0, 1, 2, 3


def con_score(cs, ce, ps, pe, K):
    cp, pp = ce - cs, pe - ps
    if cs <= ps or 5 * cp >= 9 * pp or 5 * pp >= 9 * cp:
        return NO, 0
    de, ds, dpd = abs(ce - pe), abs(cs - ps), abs(cp - pp)
    matched = min(de, K) + min(ds, K)
    v = de + ds
    lg = v.bit_length() - 1 if v else -1
    score = matched - (dpd * dpd // 2 + int(lg / 2))
    if dpd == 0:
        return (OVL if matched < 2 * K else SAME), score
    return REG, score


def chain_dp(en, pr, K, max_p, stats):
    n = len(en)
    sc = [K + min(K, p) for p in pr]
    fr = [-1] * n
    evals = 0
    nbk = max_p // S + 2
    use_buckets = nbk <= NBK_MAX and 2 * n <= 10 ** 9
    # presort: entry e = 2 i + t goes to bucket pr[i] // S - 1 + t (stable counting sort)
    if use_buckets:
        cnt = [0] * (nbk + 1)
        for i in range(n):
            b = pr[i] // S
            assert b < nbk
            cnt[b] += 1
            if b >= 1:
                cnt[b - 1] += 1
        off = [0] * (nbk + 1)
        for b in range(nbk):
            off[b + 1] = off[b] + cnt[b]
        cursor = off[:]
        idxb = [0] * off[nbk]
        for i in range(n):
            b = pr[i] // S
            for t in (0, 1):
                x = b - 1 + t
                if x >= 0:
                    idxb[cursor[x]] = i
                    cursor[x] += 1
        fill = [0] * nbk   # entries of bucket b with index < cur

        def advance(i):
            b = pr[i] // S
            fill[b] += 1
            if b >= 1:
                fill[b - 1] += 1
        advance(0)
    gmax = max(sc[0], 0) if n else 0
    for cur in range(1, n):
        ce, cp = en[cur], pr[cur]
        cs = ce - cp
        init = K + min(K, cp)
        max_score, best = init, -1
        wlo = bisect.bisect_left(en, cs, 0, cur)   # first index with en >= cs
        T = gmax + 2 * K
        Dc = int(math.sqrt(2.0 * T)) + 1           # >= the largest dpd with dpd^2 < 2 T
        fast = use_buckets and Dc <= S // 2 and (cur - wlo) < cp
        if fast:
            stats["fast"] += 1
            j = max(cp - Dc, 0) // S
            assert S * j <= max(cp - Dc, 0) and cp + Dc < S * j + 2 * S
            stop_idx = -1
            p = off[j] + fill[j] - 1
            while p >= off[j]:
                q = idxb[p]
                assert q < cur
                if en[q] < cs:
                    break
                p -= 1
                stats["cand"] += 1
                dpd = abs(cp - pr[q])
                if dpd * dpd >= 2 * T:
                    continue                        # cannot improve, cannot stop
                cls, con = con_score(cs, ce, en[q] - pr[q], en[q], K)
                if cls == NO:
                    continue
                s = sc[q] + con
                if s > max_score:
                    max_score, best = s, q
                    if cls >= SAME:
                        stop_idx = q
                        break
                elif cls == OVL:
                    stop_idx = q
                    break
            evals += cur - max(wlo, stop_idx)
        else:
            stats["slow"] += 1
            iter_n = 0
            for q in range(cur - 1, -1, -1):
                if en[q] < cs:
                    break
                evals += 1
                cls, con = con_score(cs, ce, en[q] - pr[q], en[q], K)
                if cls != NO:
                    s = sc[q] + con
                    if s > max_score:
                        max_score, best = s, q
                        if cls >= SAME:
                            break
                        iter_n = 0
                        continue
                    elif cls == OVL:
                        break
                iter_n += 1
                if iter_n >= cp:
                    break
        if max_score > init:
            sc[cur], fr[cur] = max_score, best
        gmax = max(gmax, sc[cur])
        if use_buckets:
            advance(cur)
    return sc, fr, evals


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "r2c2"
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    para = O.default_para()
    K = para.k
    stats = {"fast": 0, "slow": 0, "cand": 0}
    bad = tot_e = 0
    _, seqs = synth.gen_reads(shape, n_reads)
    for seq in seqs:
        hl = H.hits(seq, para)
        en = [h[0] for h in hl]
        pr = [h[1] for h in hl]
        if any(en[i] == en[i - 1] for i in range(1, len(en))):
            continue
        sc, fr, evals = chain_dp(en, pr, K, para.max_p, stats)
        ref = H.chain(hl, para)
        ok = ref.score == sc and ref.frm == fr and ref.n_evals == evals
        bad += not ok
        tot_e += evals
        if not ok:
            print("MISMATCH", len(hl), evals, ref.n_evals, ref.score == sc, ref.frm == fr)
    print("%s: %d reads, mismatches %d, reference evaluations %d, candidates scanned on the fast path %d (%.3f), fast %d slow %d" %
          (shape, len(seqs), bad, tot_e, stats["cand"], stats["cand"] / max(tot_e, 1), stats["fast"], stats["slow"]))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
