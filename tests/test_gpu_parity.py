"""GPU parity tests: every call goes through the C ABI (include/th_gpu.h) or the host layer above it
and is compared, bit-exact, with the CPU oracle and with the golden outputs of the unmodified
reference (tests/golden/golden.json.gz).  Integer work only => equality, no tolerances."""
import hashlib

import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import tidehunter_b200 as T
    if T.gpu_lib().th_gpu_device_count() <= 0:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box")
    return T


def _stage_inputs(golden_inputs):
    from tidehunter_b200 import synth
    names, seqs = [], []
    for tag in ("test_50x4", "test_1000x10", "testfq30"):
        n, s = golden_inputs(tag)
        names += n; seqs += s
    for shape, cnt in (("r2c2", 12), ("short", 12), ("long", 4)):
        n, s = synth.gen_reads(shape, cnt, start=100)
        names += n; seqs += s
    # edge cases: empty, shorter than k, all N, N runs inside a repeat, lower case, no repeat
    unit = b"ACGGTCATTGCAGTCCGATAGCTTAGGCTAACGTTCAGGATCCATGA"
    seqs += [b"", b"ACG", b"N" * 300, (unit + b"NNNN") * 8, (unit * 9).lower(), b"ACGTTGCAAGGCTTAACCGGTT"]
    names += [b"e%d" % i for i in range(6)]
    return names, seqs


STAGE_PARAS = [dict(), dict(k=12, w=5), dict(hpc=1), dict(min_p=2, k=5), dict(k=16), dict(hpc=1, w=4)]


@pytest.mark.parametrize("pk", range(len(STAGE_PARAS)))
def test_stage_parity(T, oracle, golden_inputs, pk):
    """hits, chain DP cells, chains and par_pos of every read, stage by stage."""
    kw = STAGE_PARAS[pk]
    names, seqs = _stage_inputs(golden_inputs)
    para = oracle.default_para(**kw)
    ctx = T.GpuContext(only_unit=1, **kw)
    ctx.process(seqs)
    bad = {"hits": [], "dp": [], "chains": [], "par": []}
    for r, s in enumerate(seqs):
        oh = H.hits(s, para)
        n, ge, gp = ctx.hits(r, max(len(s), 1))
        if n != len(oh) or list(zip(ge, gp)) != oh:
            bad["hits"].append((r, n, len(oh)))
            continue
        if len(oh) < 2:
            continue
        oc = H.chain(oh, para)
        n, gs, gf = ctx.chain_dp(r, len(oh))
        if gs != oc.score or gf != oc.frm:
            first = next(i for i in range(len(oh)) if gs[i] != oc.score[i] or gf[i] != oc.frm[i])
            bad["dp"].append((r, first, gs[first], oc.score[first], gf[first], oc.frm[first]))
            continue
        gc = ctx.chains(r, len(oh) + 8)
        if gc != oc.chains:
            bad["chains"].append((r, len(gc), len(oc.chains)))
            continue
        for ci in range(len(oc.chains)):
            op = H.partition(s, oc, ci, para)
            gpp = ctx.par_pos(r, ci)
            if gpp != op:
                bad["par"].append((r, ci, gpp[:12] if gpp else gpp, op[:12]))
        oracle.lib().tho_chain_free(oc.raw)
    ctx.close()
    assert not any(bad.values()), {k: v[:5] for k, v in bad.items() if v}


def test_ksw_global_and_extension(T, oracle):
    rng = np.random.default_rng(7)
    pairs = (H.random_pairs(rng, 60, 1, 40) + H.random_pairs(rng, 60, 30, 300, with_n=True) +
             H.random_pairs(rng, 12, 500, 1300) + H.random_pairs(rng, 3, 2500, 4200, div=0.2))
    # low complexity / tandem pairs stress the tie-breaking
    for _ in range(30):
        u = rng.integers(0, 2, int(rng.integers(1, 5)), dtype=np.uint8)
        a = np.tile(u, int(rng.integers(3, 40)))[: int(rng.integers(3, 120))]
        b = np.tile(u, int(rng.integers(3, 40)))[: int(rng.integers(3, 120))]
        pairs.append((a.tobytes(), b.tobytes()))
    qs = [p[0] for p in pairs]; ts = [p[1] for p in pairs]
    ctx = T.GpuContext()
    g0 = ctx.ksw_batch(0, qs, ts)
    exp0 = [H.ksw_global(q, t) for q, t in pairs]
    bad = [(i, len(qs[i]), len(ts[i]), g0[i][0], exp0[i][0]) for i in range(len(pairs)) if g0[i][0] != exp0[i][0]]
    assert not bad, ("global iden", bad[:8])
    # boundary projection: ksw2_global_with_cigar + ksw2_backtrack_left_end for several q_left_ext each
    q1, t1, a1, e1 = [], [], [], []
    for i, (q, t) in enumerate(pairs):
        for x in sorted({1, len(q) // 3, len(q) // 2, len(q) - 1, len(q)}):
            if 0 < x <= len(q):
                q1.append(q); t1.append(t); a1.append(x)
                e1.append((exp0[i][0], H.ksw_left_end(exp0[i][1], len(q), len(t), x)))
    g1 = ctx.ksw_batch(1, q1, t1, a1)
    bad = [(i, len(q1[i]), len(t1[i]), a1[i], g1[i], e1[i]) for i in range(len(q1)) if tuple(g1[i]) != e1[i]]
    assert not bad, ("left end", bad[:8])
    g2 = ctx.ksw_batch(2, qs, ts)
    e2 = [H.ksw_ext(q, t) for q, t in pairs]
    bad = [(i, len(qs[i]), len(ts[i]), g2[i], e2[i]) for i in range(len(pairs)) if tuple(g2[i]) != e2[i]]
    assert not bad, ("extension", bad[:8])
    # two extensions per warp (the kernel pairs extensions of similar target length): any two, no N
    q3, t3 = [], []
    for _ in range(80):
        n = int(rng.integers(1, 1400)) if rng.random() < 0.8 else int(rng.integers(1, 12))
        lowc = rng.random() < 0.3
        def mk(l):
            return (rng.integers(0, 2 if lowc else 4, l, dtype=np.uint8)).tobytes()
        qa = mk(n); qb = bytes(reversed(qa)) if rng.random() < 0.4 else mk(n if rng.random() < 0.5 else int(rng.integers(1, 1400)))
        def noisy(q, l):
            a = np.frombuffer((q * (l // len(q) + 1))[:l], dtype=np.uint8).copy()
            hit = rng.random(l) < 0.15
            a[hit] = rng.integers(0, 4, int(hit.sum()), dtype=np.uint8)
            return a.tobytes()
        q3 += [qa, qb]; t3 += [noisy(qa, int(rng.integers(0, 1500))), noisy(qb, int(rng.integers(0, 1500)))]
    g3 = ctx.ksw_batch(3, q3, t3)
    e3 = [H.ksw_ext(q, t) if len(t) else (-1, -1) for q, t in zip(q3, t3)]
    bad = [(i, len(q3[i]), len(t3[i]), g3[i], e3[i]) for i in range(len(q3)) if tuple(g3[i]) != e3[i]]
    assert not bad, ("packed extension", bad[:8])
    ctx.close()


def test_ksw_banded_pair_identity(T, oracle):
    """The packed identity alignments as ksw_pair_kernel runs them (two per warp, certified band first, th_ksw.cuh): the
    identity counts equal the oracle's full-matrix ksw2 whatever the first band width, and every path through the retry
    logic (certified at the first width, at the second, full matrix at once, full matrix after failed bands) is taken."""
    rng = np.random.default_rng(11)
    pairs = []
    for div, lo, hi, n in ((0.10, 600, 1500, 16), (0.15, 900, 1100, 24), (0.15, 1500, 2600, 8), (0.25, 700, 1400, 16), (0.20, 3000, 4500, 4),
                           (0.15, 1, 300, 24), (0.45, 600, 1200, 8)):
        pairs += H.random_pairs(rng, n, lo, hi, div=div)
    for _ in range(12):  # low complexity and internal repeats (many co-optimal paths), unrelated pairs, one long indel
        m = int(rng.integers(1, 6))
        u = np.tile(rng.integers(0, 4, m, dtype=np.uint8), 900 // m)
        pairs.append((H.random_pairs(rng, 1, 1, 1)[0][0] + u.tobytes(), u[: int(rng.integers(500, 880))].tobytes()))
        v = np.tile(rng.integers(0, 4, 97, dtype=np.uint8), 9).tobytes()
        pairs.append((v, v[int(rng.integers(0, 300)):]))
        pairs.append((rng.integers(0, 4, int(rng.integers(600, 1200)), dtype=np.uint8).tobytes(), rng.integers(0, 4, int(rng.integers(600, 1200)), dtype=np.uint8).tobytes()))
        w = rng.integers(0, 4, 1200, dtype=np.uint8)
        k = int(rng.integers(40, 400))
        pairs.append((np.concatenate([w[:500], w[500 + k:]]).tobytes(), w.tobytes()))
        pairs.append((w.tobytes(), np.concatenate([w[:300], w[300 + k:]]).tobytes()))
    order = rng.permutation(len(pairs))  # neighbours in the list are packed together: mix lengths and kinds
    pairs = [pairs[i] for i in order]
    if len(pairs) & 1:
        pairs.append(pairs[0])
    # nothing in common (every column a mismatch or two gaps): the band the failed certificate asks for is the full matrix
    pairs += [(bytes([0]) * 900, bytes([1]) * 880), (bytes([2]) * 900, bytes([3]) * 910)]
    pairs.append(pairs[1])  # odd count: the last entry runs packed with itself
    qs = [p[0] for p in pairs]; ts = [p[1] for p in pairs]
    exp = [H.ksw_global(q, t)[0] for q, t in pairs]
    ctx = T.GpuContext()
    paths = set()
    for alpha in (0, 20, 60, 120, 250, 400):
        g = ctx.ksw_batch(4, qs, ts, [alpha] * len(pairs))
        bad = [(alpha, i, len(qs[i]), len(ts[i]), g[i], exp[i]) for i in range(len(pairs)) if g[i][0] != exp[i]]
        assert not bad, ("packed banded identity", bad[:8])
        paths |= {x[1] for x in g}
    ctx.close()
    assert {0, 1, 2}.issubset(paths) and (3 in paths or 4 in paths), paths


def _run_case(T, golden_inputs, c, **extra):
    names, seqs = golden_inputs(c["input"])
    th = T.TideHunter(**dict(c["para"], **extra))
    out = th.run(names, seqs)
    st = th.stats()
    th.close()
    return out, st


def test_config1_golden(T, golden, golden_inputs):
    """BASELINE config 1: test_data/test_1000x10.fa, default options, FASTA -- byte-identical."""
    c = next(c for c in golden["cases"] if c["input"] == "test_1000x10" and c["args"] == ["-f", "1"])
    out, st = _run_case(T, golden_inputs, c)
    assert out.decode() == c["text"]
    assert hashlib.md5(out).hexdigest() == "6518be7cff7c168de0e4c42465b3a6dd"
    assert st["n_launches"] > 0 and st["n_poa_cells"] > 0 and st["n_ksw_cells"] > 0


def test_all_golden_cases(T, golden, golden_inputs):
    bad = []
    for c in golden["cases"]:
        out, _ = _run_case(T, golden_inputs, c)
        if hashlib.md5(out).hexdigest() != c["md5"]:
            exp = c.get("text", "")
            got = out.decode()
            gl, el = got.split("\n"), exp.split("\n")
            first = next((i for i in range(min(len(gl), len(el))) if gl[i] != el[i]), min(len(gl), len(el)))
            bad.append((c["input"], c["args"], len(gl), len(el), first, gl[first][:160] if first < len(gl) else None, el[first][:160] if first < len(el) else None))
    assert not bad, bad[:6]


def test_parallel_formatting_does_not_change_output(T, golden, golden_inputs, monkeypatch):
    """th_host formats a finished chunk on several threads (segments of reads, own buffers, concatenated in order);
    with a grain of 3 reads per thread the 100-read FASTQ cases (quality slot quirk included) and the adapter cases
    run through that path."""
    monkeypatch.setenv("TH_HOST_FMT_GRAIN", "3")
    bad = []
    for c in golden["cases"]:
        if c["input"] not in ("testfq_all", "full_length"):
            continue
        out, _ = _run_case(T, golden_inputs, c)
        if hashlib.md5(out).hexdigest() != c["md5"]:
            bad.append((c["input"], c["args"]))
    assert not bad, bad


@pytest.mark.parametrize("lanes", [1, 2, 4])
def test_chunking_does_not_change_output(T, golden, golden_inputs, lanes):
    """Chunks rotate over `lanes` GPU contexts on their own host threads (host/th_host.c); the text, including the
    FASTQ quality slot quirk that depends on the global read order, must not depend on chunk size or lane count."""
    c = next(c for c in golden["cases"] if c["input"] == "testfq_all" and c["args"] == ["-f", "4"])
    out, st = _run_case(T, golden_inputs, c, chunk_reads=7, lanes=lanes)
    assert hashlib.md5(out).hexdigest() == c["md5"]
    assert st["n_launches"] >= 10 * ((len(golden_inputs(c["input"])[0]) + 6) // 7)


@pytest.mark.parametrize("shape,n", [("r2c2", 96), ("short", 128), ("long", 24)])
def test_synthetic_vs_oracle(T, oracle, shape, n):
    from tidehunter_b200 import synth
    names, seqs = synth.gen_reads(shape, n, start=1000)
    exp, cnt = oracle.run_batch(names, seqs, oracle.default_para(out_fmt=2), threads=8)
    th = T.TideHunter(out_fmt=2)
    out = th.run(names, seqs)
    st = th.stats()
    th.close()
    if out != exp:
        gl, el = out.decode().split("\n"), exp.decode().split("\n")
        diff = [(i, gl[i][:150], el[i][:150]) for i in range(min(len(gl), len(el))) if gl[i] != el[i]]
        pytest.fail("%d/%d lines differ (%d vs %d lines); first: %r" % (len(diff), len(el), len(gl), len(el), diff[:2]))
    # the GPU's work counters are the roofline numerators: they must equal the oracle's
    assert st["n_hits"] == cnt["hits"]
    assert st["n_chain_evals"] == cnt["chain_evals"]
    assert st["n_poa_cells"] == cnt["poa_cells"]


def test_edge_reads(T, oracle):
    names = [b"e", b"s", b"n", b"u"]
    seqs = [b"", b"ACGT", b"N" * 500, b"ACGTTGCA" * 4 + b"GATTACAGATTACCA"]
    th = T.TideHunter()
    assert th.run(names, seqs) == oracle.run_batch(names, seqs, oracle.default_para())[0]
    assert th.run([], []) == b""
    th.close()


def test_cli_golden_cases(T, golden, golden_inputs, tmp_path):
    """The command line front end (host/tidehunter-b200: reader thread -> host layer -> C ABI) on files: every golden
    case, with the input written as multi-line FASTA, as gzip, and (quality-format cases) as FASTQ."""
    import gzip
    import os
    import subprocess
    from tidehunter_b200 import build as B
    cli = B.build()[2]
    five, three = golden["adapters"]["five"], golden["adapters"]["three"]
    p5, p3 = str(tmp_path / "5.fa"), str(tmp_path / "3.fa")
    open(p5, "w").write(">5\n%s\n" % five)
    open(p3, "w").write(">3 adapter\n%s\n%s\n" % (three[:10], three[10:]))
    bad = []
    for k, c in enumerate(golden["cases"]):
        names, seqs = golden_inputs(c["input"])
        style = k % 3
        path = str(tmp_path / ("in%d.fx" % k))
        with open(path, "wb") as f:
            for n, s in zip(names, seqs):
                if style == 1:      # FASTQ, quality may start with '@'
                    f.write(b"@" + n + b" some comment\n" + s + b"\n+" + n + b"\n" + b"@" * len(s) + b"\n")
                else:               # FASTA, 70 columns, CRLF for every third case
                    eol = b"\r\n" if k % 9 == 0 else b"\n"
                    f.write(b">" + n + b"\tcomment" + eol + eol.join(s[i:i + 70] for i in range(0, len(s), 70)) + eol)
        if style == 2:
            with open(path, "rb") as f, gzip.open(path + ".gz", "wb") as g:
                g.write(f.read())
            path += ".gz"
        args = []
        i = 0
        while i < len(c["args"]):   # golden args name the reference's adapter files: point them at ours
            a = c["args"][i]
            if a in ("-5", "-3"):
                args += [a, p5 if a == "-5" else p3]; i += 2
            else:
                args.append(a); i += 1
        r = subprocess.run([cli] + args + [path], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if r.returncode != 0 or hashlib.md5(r.stdout).hexdigest() != c["md5"]:
            bad.append((c["input"], c["args"], style, r.returncode, r.stderr.decode()[-200:]))
    assert not bad, bad[:5]


def test_baseline_workload_chunk_matches_reference_md5(T):
    """One 16,384-read chunk of the full BASELINE configs[1] workload (1,048,576 reads) against the md5 of the unmodified
    reference's output for it (tests/golden/r2c2_1m_md5.json, made by tools/million_parity.py --make-md5).  The
    whole set is checked by tools/million_parity.py --check (profiles/r1_million_parity.json)."""
    import json
    import os
    from tidehunter_b200 import synth
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "r2c2_1m_md5.json")))
    assert len(fx["chunks"]) == 64 and sum(c["reads"] for c in fx["chunks"]) == 1048576
    c = fx["chunks"][37]
    names, seqs = synth.gen_reads("r2c2", c["reads"], start=c["first_read"])
    assert hashlib.md5(b"".join(seqs)).hexdigest() == c["input_md5"]
    th = T.TideHunter(out_fmt=1)
    out = th.run(names, seqs)
    th.close()
    assert len(out) == c["bytes"] and hashlib.md5(out).hexdigest() == c["md5"]


def test_sse_vector_width_mode(T):
    """simd_lanes16 = 8: the GPU path emulating abPOA's SSE4.1 build (template instance LP = 3 of the POA kernel)
    against the output of the reference built that way (tests/golden/pn8_golden.json)."""
    import json
    import os
    from tidehunter_b200 import synth
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pn8_golden.json")))
    names, seqs = [], []
    for shape, start, n in fx["sets"]:
        a, b = synth.gen_reads(shape, n, start=start)
        names += a; seqs += b
    for lanes16, md5 in ((8, fx["md5_pn8"]), (16, fx["md5_pn16"])):
        th = T.TideHunter(out_fmt=2, simd_lanes16=lanes16)
        out = th.run(names, seqs)
        th.close()
        assert hashlib.md5(out).hexdigest() == md5, "simd_lanes16 = %d" % lanes16


def test_one_process_several_devices(T, golden, golden_inputs):
    """th_host_create_multi: lanes spread over a device list (here every visible GPU, and device 0 listed twice so the
    path is exercised on a one-GPU box too); output must not depend on it."""
    c = next(c for c in golden["cases"] if c["input"] == "testfq_all" and c["args"] == ["-f", "4"])
    names, seqs = golden_inputs(c["input"])
    n_dev = T.gpu_lib().th_gpu_device_count()
    for devs in ([0, 0], list(range(n_dev))):
        th = T.TideHunter(devices=devs, lanes=2, chunk_reads=7, **c["para"])
        out = th.run(names, seqs)
        th.close()
        assert hashlib.md5(out).hexdigest() == c["md5"], devs


def test_gap_modes(T):
    """Convex (default), affine (-O x,0) and linear (-O 0,...) abPOA gap modes on long-indel reads against the reference's
    outputs (tests/golden/gapmode_golden.json).  The linear mode runs every consensus on the wide pass."""
    import json
    import os
    from tidehunter_b200 import synth
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gapmode_golden.json")))
    names, seqs = synth.gen_long_indel_reads(fx["n_reads"])
    for tag, m in list(fx["modes"].items()) + list(fx["oracle_only_modes"].items()):
        th = T.TideHunter(out_fmt=2, **m["para"])
        out = th.run(names, seqs)
        failed = th.failed_tasks()
        th.close()
        assert failed == 0, tag
        assert hashlib.md5(out).hexdigest() == m["md5"], tag


def test_very_long_reads(T, oracle):
    """~100 kb reads: 60 and 800 copies in one consensus task, 6 kb units -- capacities (sorts beyond shared memory, slab
    sizing with the full-width retry, sink in-degree, cigar lengths), not throughput."""
    from tidehunter_b200 import synth
    names, seqs = synth.gen_very_long_reads()
    exp, cnt = oracle.run_batch(names, seqs, oracle.default_para(out_fmt=2), threads=3)
    assert exp.count(b"\n") >= 3
    th = T.TideHunter(out_fmt=2)
    out = th.run(names, seqs)
    st = th.stats()
    th.close()
    assert out == exp
    assert st["n_poa_cells"] == cnt["poa_cells"] and st["n_chain_evals"] == cnt["chain_evals"]


def test_int32_range_units(T, oracle):
    """A unit whose graph leaves abPOA's int16 score range (synth.gen_int32_read: the reference switches to 32-bit vectors
    of half the lanes for the last alignments, simd_abpoa_align.c:1610-1621).  The packed 16-bit kernel hands the task to
    the wide pass, which picks the score width per alignment as the reference does: the record is the reference's
    (tests/golden/int32_golden.json), nothing is dropped, and the reads around it are unaffected."""
    import hashlib, json
    from tidehunter_b200 import synth
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "int32_golden.json")))
    n0, s0 = synth.gen_reads("r2c2", 3, start=42)
    nw, sw = synth.gen_int32_read()
    th = T.TideHunter(out_fmt=2)
    ref3 = th.run(n0, s0)
    solo = th.run(nw, sw)
    assert hashlib.md5(solo).hexdigest() == gold["md5"] and len(solo) == int(gold["bytes"])
    out = th.run(n0[:2] + nw + n0[2:], s0[:2] + sw + s0[2:])
    failed = th.failed_tasks()
    th.close()
    assert failed == 0
    def of(read, text):
        return [l for l in text.split(b"\n") if l.split(b"\t")[0] == read]
    n0b = [x if isinstance(x, bytes) else x.encode() for x in n0]
    exp = of(n0b[0], ref3) + of(n0b[1], ref3) + of(b"w0", solo) + of(n0b[2], ref3)
    assert [l for l in out.split(b"\n") if l] == exp and of(b"w0", solo)


def test_option_sets_found_by_fuzzing(T):
    """The same option sets (tests/golden/fuzz_found_golden.json: -l with quality formats) through the host layer."""
    import json
    import os
    from tidehunter_b200 import synth
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fuzz_found_golden.json")))["cases"]
    for c in fx:
        names, seqs = synth.gen_reads(c["shape"], c["n"], start=c["start"])
        th = T.TideHunter(**c["para"])
        out = th.run(names, seqs)
        th.close()
        assert hashlib.md5(out).hexdigest() == c["md5"], c["args"]


def test_sharded_parts_concatenate_to_the_whole(T):
    """Two processes' parts of one input, each through its own object (different chunking and lanes, one through a Batch and
    the zero-copy view): the outputs concatenate to the single-process output -- for FASTQ as long as the input has at most
    4,096 reads (beyond that the reference prints stale quality bytes of the read 4,096 positions earlier from a buffer it
    never rewinds, src/main.c:262-266, which a part that starts elsewhere cannot know), for the tabular format always."""
    from tidehunter_b200 import synth
    for fmt, n, cut in ((3, 3000, 1700), (2, 4600, 4300)):
        names, seqs = synth.gen_reads("short", n, start=31000)
        th = T.TideHunter(out_fmt=fmt, chunk_reads=1024, lanes=2)
        whole = th.run(names, seqs)
        th.close()
        a = T.TideHunter(out_fmt=fmt, chunk_reads=1024, lanes=2)
        b = T.TideHunter(out_fmt=fmt, chunk_reads=700, lanes=3)
        pa = a.run(names[:cut], seqs[:cut])
        pb = bytes(b.run(T.Batch(names[cut:], seqs[cut:]), first_index=cut, copy=False))
        a.close(); b.close()
        assert pa + pb == whole, fmt
        assert whole.count(b"\n") > n
