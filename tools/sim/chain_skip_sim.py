"""Chaining DP restated in Python (tandem_chain.c:290-356 as th_chain.cuh runs it: one hit per end, predecessors in batches of 32)
with a candidate pruning rule counted on top: a 32-aligned batch of predecessors is skipped when its best score plus 2k cannot
beat the running maximum and it lies further than k bases back (no overlap stop possible).  Prints how many batches the rule
would skip; scores, links and the evaluation counter are checked against the oracle.  Development aid.
usage: python tools/sim/chain_skip_sim.py [shape] [reads]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for d in ('', 'oracle', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, d))
import numpy as np
import helpers as H
import oracle_py as O
from tidehunter_b200 import synth
K = 8
def con_score(cs, ce, ps, pe):
    cp, pp = ce - cs, pe - ps
    if cs <= ps or 5 * cp >= 9 * pp or 5 * pp >= 9 * cp: return 0, 0
    de, ds, dpd = abs(ce - pe), abs(cs - ps), abs(cp - pp)
    matched = min(de, K) + min(ds, K)
    v = de + ds
    lg = v.bit_length() - 1 if v else -1
    score = matched - (dpd * dpd // 2 + int(lg / 2))
    if dpd == 0: return (3 if matched < 2 * K else 2), score
    return 1, score
shape = sys.argv[1] if len(sys.argv) > 1 else "r2c2"
names, seqs = synth.gen_reads(shape, int(sys.argv[2]) if len(sys.argv) > 2 else 2)
para = O.default_para()
tot_b = skip_b = tot_e = 0
for seq in seqs:
    hl = H.hits(seq, para)
    n = len(hl)
    en = [h[0] for h in hl]; pr = [h[1] for h in hl]
    if any(en[i] == en[i-1] for i in range(1, n)): print("dup ends, skip"); continue
    sc = [K + min(K, p) for p in pr]; fr = [-1] * n
    nb = (n + 31) // 32
    bmax = [0] * nb
    evals = 0
    for cur in range(1, n):
        if cur % 32 == 0: bmax[cur // 32 - 1] = max(sc[cur - 32:cur])
        ce, cp = en[cur], pr[cur]; cs = ce - cp
        init = K + min(K, cp); max_score = init; best = -1; iter_n = 0; max_h = cp
        pre = cur - 1; stop = False
        first_block = (cur - 1) // 32
        while pre >= 0 and not stop:
            b = pre // 32
            lo = 32 * b
            # batch = hits [lo, pre]
            tot_b += 1
            if b < first_block and pre == lo + 31 and en[pre] <= ce - K and bmax[b] + 2 * K <= max_score:
                # skippable: nothing improves, no OVL stop; account the rows
                skip_b += 1
                cnt = 0
                for q in range(pre, lo - 1, -1):
                    if en[q] < cs: stop = True; break
                    cnt += 1; iter_n += 1
                    # verify the claim
                    cls, con = con_score(cs, ce, en[q] - pr[q], en[q])
                    assert not (cls and sc[q] + con > max_score), "skip rule violated (improve)"
                    assert cls != 3, "skip rule violated (OVL)"
                    if iter_n >= max_h: stop = True; break
                evals += cnt
                pre = lo - 1
                continue
            for q in range(pre, lo - 1, -1):
                if en[q] < cs: stop = True; break
                evals += 1
                cls, con = con_score(cs, ce, en[q] - pr[q], en[q])
                if cls:
                    s = sc[q] + con
                    if s > max_score:
                        max_score = s; best = q
                        if cls >= 2: stop = True; break
                        iter_n = 0; continue
                    elif cls == 3: stop = True; break
                iter_n += 1
                if iter_n >= max_h: stop = True; break
            pre = lo - 1
        if max_score > init: sc[cur] = max_score; fr[cur] = best
    ref = H.chain(hl, para)
    ok = (ref.score == sc and ref.frm == fr)
    print("hits", n, "evals", evals, "oracle evals", ref.n_evals, "dp equal", ok, "batches", tot_b, "skipped", skip_b)
    tot_e += evals
print("skipped share of batches: %.3f" % (skip_b / max(tot_b, 1)))
