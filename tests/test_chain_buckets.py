"""The period-bucketed chaining DP (tidehunter_b200/csrc/th_chain.cuh) on the CPU: tools/sim/chain_bucket_sim.py models the
algorithm (which predecessors may be skipped, how the reference's evaluation count follows from indices, when the plain scan
is used) and tools/sim/chain_warp_emu.py emulates the kernel's warp code in lock step (stable scatter with __match_any_sync
ranks, 32-ary search, ballots / shuffles of the batch logic).  Both must reproduce the oracle's tandem_chain
(src/tandem_chain.c:290-356) in scores, links and evaluation count.  The kernel itself: tests/test_gpu_parity.py::test_stage_parity."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "sim"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H  # noqa: E402
from tidehunter_b200 import synth  # noqa: E402
import chain_bucket_sim as M  # noqa: E402
import chain_warp_emu as E  # noqa: E402

O = H.O   # the oracle binding the helpers use (one module object: its ctypes classes must match)


def _random_hit_list(rng, trial):
    mode = trial % 5
    L = int(rng.integers(200, 3000))
    n = int(rng.integers(2, min(L // 2 - 1, 300)))
    ends = np.sort(rng.choice(np.arange(10, L), size=n, replace=False))
    if mode == 0:
        pers = rng.integers(2, 400, n)
    elif mode == 1:
        pers = rng.choice([100, 101, 99, 200, 300, 50], n)          # many equal periods: the overlap / same-period stops
    elif mode == 2:
        pers = rng.integers(90, 110, n)
    elif mode == 3:
        pers = np.where(rng.random(n) < 0.5, 1000 + rng.integers(-3, 4, n), rng.integers(30, 5000, n))
    else:
        pers = rng.integers(500, 530, n)                            # straddles a bucket border
    pers = np.maximum(np.minimum(pers, ends), 1)
    return L, ends, pers


@pytest.mark.parametrize("engine", ["model", "warp"])
def test_bucketed_chain_dp_equals_oracle(engine):
    rng = np.random.default_rng(23)
    n_fast = 0
    for trial in range(60 if engine == "warp" else 200):
        L, ends, pers = _random_hit_list(rng, trial)
        para = O.default_para(k=int(rng.choice([5, 8, 8, 12])), max_p=int(rng.choice([10000, 10000, 600, 30000] + ([40000] if engine == "model" else []))))
        pers = np.minimum(pers, para.max_p)
        hl = [(int(e), int(p)) for e, p in zip(ends, pers)]
        en, pr = [h[0] for h in hl], [h[1] for h in hl]
        if engine == "model":
            st = {"fast": 0, "slow": 0, "cand": 0}
            sc, fr, ev = M.chain_dp(en, pr, para.k, para.max_p, st)
            n_fast += st["fast"]
        else:
            sc, fr, ev, nf = E.chain_read(en, pr, para.k, para.max_p, (L + 63) // 64 * 64)
            n_fast += nf
        ref = H.chain(hl, para)
        assert ref.score == sc and ref.frm == fr and ref.n_evals == ev, (engine, trial, len(hl))
    assert n_fast > 1000


def test_bucketed_chain_dp_on_reads():
    para = O.default_para()
    for shape in ("r2c2", "mixed"):
        _, seqs = synth.gen_reads(shape, 2)
        for seq in seqs:
            hl = H.hits(seq, para)
            en, pr = [h[0] for h in hl], [h[1] for h in hl]
            if len(hl) < 2 or any(en[i] == en[i - 1] for i in range(1, len(en))):
                continue
            st = {"fast": 0, "slow": 0, "cand": 0}
            sc, fr, ev = M.chain_dp(en, pr, para.k, para.max_p, st)
            ref = H.chain(hl, para)
            assert ref.score == sc and ref.frm == fr and ref.n_evals == ev
            assert st["cand"] <= ev
            if shape == "r2c2":
                assert st["cand"] * 3 < ev  # the point of it: far fewer predecessors are looked at than the reference evaluates
