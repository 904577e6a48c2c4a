"""Seeded synthetic tandem-repeat read generator for the BASELINE.json workloads.

Each read = 50 bp random flank + `copies` copies of a uniform-random unit (rotated by a random
offset) + 50 bp random flank, passed through a per-base error channel: with probability `err` a
template base is hit by an error, split 40 % substitution / 30 % insertion (a uniform base is
inserted before the template base) / 30 % deletion.  The PRNG is numpy's counter-based Philox keyed by
(seed, shape id), so every rank / test regenerates the same reads without sharing state.

Shapes (BASELINE.json `configs`):
  r2c2   : unit 1000, 10 copies, err 0.15            (config 2, the bench workload)
  short  : unit U{50..200}, copies U{20..50}, err 0.10  (config 3)
  long   : unit U{4000..5000}, copies U{2..4}, err 0.20  (config 4)
  mixed  : read length drawn from test_data/test.fq's 100 lengths (1.8 - 23.6 kb), unit U{800..2400} (test.fq's consensus lengths), err 0.12
"""
import numpy as np

SEED = 20260117
SHAPES = {
    "r2c2": dict(unit=(1000, 1000), copies=(10, 10), err=0.15, sid=0),
    "short": dict(unit=(50, 200), copies=(20, 50), err=0.10, sid=1),
    "long": dict(unit=(4000, 5000), copies=(2, 4), err=0.20, sid=2),
}
# read lengths of the reference's test_data/test.fq (100 ONT reads, 1.8 - 23.6 kb): the "mixed" shape draws its read lengths
# from this list, so a batch has the length mix of real data -- the case where dealing equal read counts is unequal work
TESTFQ_LENGTHS = [1813, 1885, 1894, 2012, 2028, 2031, 2080, 2092, 2104, 2114, 2140, 2162, 2181, 2209, 2249, 2258, 2264, 2313, 2337, 2341,
                  2368, 2373, 2503, 2508, 2509, 2539, 2556, 2607, 2618, 2638, 2697, 2714, 2747, 2783, 2798, 2805, 2890, 2904, 2921, 3011,
                  3023, 3026, 3029, 3116, 3127, 3134, 3136, 3161, 3182, 3191, 3192, 3194, 3221, 3238, 3254, 3278, 3311, 3317, 3348, 3390,
                  3411, 3465, 3478, 3538, 3560, 3607, 3745, 3805, 4131, 4183, 4464, 4597, 4640, 4717, 4745, 4868, 5197, 5208, 5211, 5231,
                  5269, 5326, 5572, 5696, 6042, 6044, 6089, 6269, 6463, 6807, 6943, 7152, 7461, 7535, 7604, 7861, 8471, 9390, 14329, 23611]
SHAPES["mixed"] = dict(unit=(800, 2400), lengths=TESTFQ_LENGTHS, err=0.12, sid=3)   # units: the range of test.fq's consensus lengths (880 - 2416)
SHAPES["mixed23k"] = dict(unit=(800, 2400), lengths=[23611], err=0.12, sid=4)    # the two ends of the mix, for per-stage profiles
SHAPES["mixed2k"] = dict(unit=(800, 2400), lengths=TESTFQ_LENGTHS[:10], err=0.12, sid=5)
FLANK = 50


def _unit_and_copies(cfg, rng):
    """The first draws of a read: unit length and copy number (the "mixed" shape: a read length from the list, a unit
    length, and as many copies as fill the read)."""
    if "lengths" in cfg:
        length = int(cfg["lengths"][int(rng.integers(0, len(cfg["lengths"])))])
        ulen = int(rng.integers(cfg["unit"][0], cfg["unit"][1] + 1))
        return ulen, max(2, int(round((length - 2 * FLANK) / ulen)))
    ulen = int(rng.integers(cfg["unit"][0], cfg["unit"][1] + 1))
    copies = int(rng.integers(cfg["copies"][0], cfg["copies"][1] + 1))
    return ulen, copies


def nominal_lengths(shape, n, start=0, seed=SEED):
    """Template lengths (before the error channel) of reads start..start+n-1 without generating them: what a sharded run
    deals its reads by."""
    cfg = SHAPES[shape]
    out = np.empty(n, dtype=np.int64)
    for k, i in enumerate(range(start, start + n)):
        rng = np.random.Generator(np.random.Philox(key=[seed + cfg["sid"], i]))
        ulen, copies = _unit_and_copies(cfg, rng)
        out[k] = 2 * FLANK + ulen * copies
    return out
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def gen_reads(shape, n, start=0, seed=SEED, adapters=None, three_rc=False):
    """Return (names, seqs): `n` reads of `shape`, read indices start..start+n-1.

    Read i depends only on (seed, shape, i).  `adapters` = (five, three) splices
    five + unit + revcomp(three)-style R2C2 structure: the unit itself becomes
    three_rc... kept simple: unit' = five + unit + three (both as given), so -5/-3 searches succeed.
    """
    return gen_reads_at(shape, range(start, start + n), seed=seed, adapters=adapters, three_rc=three_rc)


def gen_reads_at(shape, indices, seed=SEED, adapters=None, three_rc=False):
    """gen_reads for an arbitrary list of read indices (a rank's part of a length-sorted batch)."""
    cfg = SHAPES[shape]
    names, seqs = [], []
    for i in indices:
        i = int(i)
        rng = np.random.Generator(np.random.Philox(key=[seed + cfg["sid"], i]))
        ulen, copies = _unit_and_copies(cfg, rng)
        unit = rng.integers(0, 4, ulen, dtype=np.uint8)
        if adapters is not None:
            five, three = adapters
            a5 = np.frombuffer(five.encode(), dtype=np.uint8)
            a3 = np.frombuffer(three.encode(), dtype=np.uint8)
            lut = np.zeros(256, dtype=np.uint8)
            lut[ord("A")], lut[ord("C")], lut[ord("G")], lut[ord("T")] = 0, 1, 2, 3
            # three_rc: the splint structure -F looks for (5' adapter, insert, reverse complement of the 3' adapter:
            # src/gen_cons.c:224-291 searches five_seq and revcomp(three_seq) on the same strand)
            unit = np.concatenate([lut[a5], unit, (3 - lut[a3])[::-1] if three_rc else lut[a3]])
            ulen = len(unit)
        rot = int(rng.integers(0, ulen))
        body = np.tile(unit, copies + 1)[rot:rot + ulen * copies]
        tmpl = np.concatenate([rng.integers(0, 4, FLANK, dtype=np.uint8), body, rng.integers(0, 4, FLANK, dtype=np.uint8)])
        u = rng.random(len(tmpl))
        e = cfg["err"]
        sub = u < 0.4 * e
        ins = (u >= 0.4 * e) & (u < 0.7 * e)
        dele = (u >= 0.7 * e) & (u < e)
        rb = rng.integers(0, 4, len(tmpl), dtype=np.uint8)   # replacement / inserted bases
        shift = rng.integers(1, 4, len(tmpl), dtype=np.uint8)
        base = np.where(sub, (tmpl + shift) & 3, tmpl)
        cnt = np.where(dele, 0, np.where(ins, 2, 1))
        out = np.repeat(base, cnt)
        # first emitted base of an insertion site is the random inserted base
        pos = np.cumsum(cnt) - cnt
        out[pos[ins]] = rb[ins]
        seqs.append(_ACGT[out].tobytes())
        names.append(("%s%d" % (shape[0], i)).encode())
    return names, seqs


def _channel(rng, tmpl, e):
    """The error channel of gen_reads on a template of codes 0..3."""
    u = rng.random(len(tmpl))
    sub = u < 0.4 * e
    ins = (u >= 0.4 * e) & (u < 0.7 * e)
    dele = (u >= 0.7 * e) & (u < e)
    rb = rng.integers(0, 4, len(tmpl), dtype=np.uint8)
    shift = rng.integers(1, 4, len(tmpl), dtype=np.uint8)
    base = np.where(sub, (tmpl + shift) & 3, tmpl)
    cnt = np.where(dele, 0, np.where(ins, 2, 1))
    out = np.repeat(base, cnt)
    pos = np.cumsum(cnt) - cnt
    out[pos[ins]] = rb[ins]
    return out


def gen_single_copy(n, adapters, start=0, seed=SEED, err=0.08):
    """Reads for the -s (single-copy full-length) path, src/gen_cons.c:128-171: flank + 5' adapter + one insert of
    U{300..2500} bp + reverse complement of the 3' adapter + flank, every other read reverse-complemented as a whole,
    through the same error channel (err lower than the R2C2 shape so that most adapters stay above -a 0.8).  Every fourth
    read carries the insert twice in tandem, so chains and the single-copy scan both produce records."""
    five, three = adapters
    lut = np.zeros(256, dtype=np.uint8)
    lut[ord("A")], lut[ord("C")], lut[ord("G")], lut[ord("T")] = 0, 1, 2, 3
    a5 = lut[np.frombuffer(five.encode(), dtype=np.uint8)]
    a3rc = (3 - lut[np.frombuffer(three.encode(), dtype=np.uint8)])[::-1]
    names, seqs = [], []
    for i in range(start, start + n):
        rng = np.random.Generator(np.random.Philox(key=[seed + 7, i]))
        ins_len = int(rng.integers(300, 2501))
        insert = rng.integers(0, 4, ins_len, dtype=np.uint8)
        body = np.concatenate([a5, insert, a3rc])
        if i % 4 == 3:
            body = np.concatenate([body, body, body[:len(body) // 2]])
        tmpl = np.concatenate([rng.integers(0, 4, FLANK, dtype=np.uint8), body, rng.integers(0, 4, FLANK, dtype=np.uint8)])
        if i % 2 == 1:
            tmpl = (3 - tmpl)[::-1]
        out = _channel(rng, tmpl, err)
        seqs.append(_ACGT[out].tobytes())
        names.append(("s%d" % i).encode())
    return names, seqs


def total_bases(seqs):
    return int(sum(len(s) for s in seqs))


def gen_long_indel_reads(n, start=0, seed=SEED + 7, err=0.08):
    """Tandem repeats (unit 300-900 bp, 4-8 copies) whose copies carry up to two 22-59 bp deletions or insertions on top
    of the per-base error channel: gaps long enough for the second gap function of abPOA's convex model (O2 = 24, E2 = 1
    beats O1 = 4, E1 = 2 beyond 20 columns) to decide the alignment, which the BASELINE shapes (single-base errors) never
    do.  Used for the convex-vs-affine gap mode tests.  Read i depends only on (seed, i)."""
    names, seqs = [], []
    for i in range(start, start + n):
        rng = np.random.Generator(np.random.Philox(key=[seed, i]))
        ulen = int(rng.integers(300, 900))
        copies = int(rng.integers(4, 9))
        unit = rng.integers(0, 4, ulen, dtype=np.uint8)
        parts = [rng.integers(0, 4, 40, dtype=np.uint8)]
        for _ in range(copies):
            u = unit.copy()
            for _ in range(int(rng.integers(0, 3))):
                gl = int(rng.integers(22, 60))
                pos = int(rng.integers(10, len(u) - gl - 10))
                if rng.random() < 0.5:
                    u = np.concatenate([u[:pos], u[pos + gl:]])
                else:
                    u = np.concatenate([u[:pos], rng.integers(0, 4, gl, dtype=np.uint8), u[pos:]])
            parts.append(u)
        parts.append(rng.integers(0, 4, 40, dtype=np.uint8))
        out = _channel(rng, np.concatenate(parts), err)
        seqs.append(_ACGT[out].tobytes())
        names.append(("g%d" % i).encode())
    return names, seqs


def gen_very_long_reads(seed=SEED + 11):
    """Three ~100 kb reads that stress capacities rather than throughput: 2 kb x 60 copies, 150 bp x 800 copies (a
    consensus task with 800 sequences), 6 kb x 12 copies (units close to the int16 score range of abPOA)."""
    names, seqs = [], []
    for i, (ulen, copies) in enumerate(((2000, 60), (150, 800), (6000, 12))):
        rng = np.random.Generator(np.random.Philox(key=[seed, i]))
        unit = rng.integers(0, 4, ulen, dtype=np.uint8)
        tmpl = np.concatenate([rng.integers(0, 4, FLANK, dtype=np.uint8), np.tile(unit, copies), rng.integers(0, 4, FLANK, dtype=np.uint8)])
        seqs.append(_ACGT[_channel(rng, tmpl, 0.10)].tobytes())
        names.append(("vl%d" % i).encode())
    return names, seqs


def gen_int32_read(seed=SEED + 13):
    """One 98 kb read, unit 9.8 kb x 10 copies at 15 % error: the partial-order graph grows beyond 16.4 k rows, where
    abPOA leaves its int16 score range and switches to 32-bit vectors (simd_abpoa_align.c:1610-1621) for the last
    alignments.  The oracle follows (pinned on the reference's output); the GPU path reports the task as failed."""
    rng = np.random.Generator(np.random.Philox(key=[seed, 0]))
    unit = rng.integers(0, 4, 9800, dtype=np.uint8)
    tmpl = np.concatenate([rng.integers(0, 4, FLANK, dtype=np.uint8), np.tile(unit, 10), rng.integers(0, 4, FLANK, dtype=np.uint8)])
    return [b"w0"], [_ACGT[_channel(rng, tmpl, 0.15)].tobytes()]
