#!/usr/bin/env python
"""bench.py -- throughput of the TideHunter per-read hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): synthetic ONT R2C2-style reads, 10 kb, 1 kb unit x 10 copies, 15 %
error, default options (-k 8 -w 1 -p 30 -P 10000 -c 2 -e 0.25, -f 1).  A step = one pass of the whole hot
path (pack, seeding, chaining, partition, POA consensus, ksw2 identity/extension) over one batch of
`--reads` reads PER GPU (weak scaling; reads are independent, no data-path collective).

  value : reads/s over all ranks, reads already resident in HBM (th_gpu_upload before the timed
          region, th_gpu_process_resident per step); device time from CUDA events on the lanes' own
          streams (th_gpu_mark: latest end - earliest start), max over ranks.
  e2e   : the same metric through the user-facing call (host layer th_host_run over the C ABI) with
          HOST buffers: pinned staging + H2D, all kernels, D2H of the results, record formatting, and
          the host-side ordered gather of the output text on rank 0.
  roofline / kernels : per-stage device time from CUDA events recorded on the library's own stream
          (th_gpu_stats), algorithmic work counters from the kernels themselves.
  cpu_baseline : the unmodified reference (oracle/_ref/TideHunter, -t <all cores>) on a bounded
          sample of the same reads (rank 0, N = 1 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Per-kernel constants taken from the ncu --set full captures under profiles/ (tools/ncu_constants.py writes the file):
# warp-instructions, ALU-pipe warp-instructions and DRAM bytes per algorithmic unit (POA / ksw cell, chain pair evaluation).
# bench.py times the kernels live (CUDA events); these constants turn that time into pipe utilisation and DRAM traffic.
def load_ncu_constants():
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_constants.json")) as f:
            return json.load(f)
    except Exception:
        return {}


WORKLOADS = {
    "r2c2": "synthetic ONT R2C2-style reads: 10 kb, 1 kb unit x 10 copies, 15% error (BASELINE.json configs[1])",
    "mixed": "synthetic reads with the length mix of test_data/test.fq (1.8-23.6 kb, unit 800-2400 bp, 12% error), input order random",
    "mixed_sorted": "synthetic reads with the length mix of test_data/test.fq (1.8-23.6 kb, unit 800-2400 bp, 12% error), input sorted by length",
}
WORKLOAD = WORKLOADS["r2c2"]


def make_config(args, world):
    """The `config` object of the JSON line -- the same for both arms (the reference arm times a bounded sample of it)."""
    return {"workload": WORKLOADS[args.workload], "reads_per_gpu_per_step": args.reads, "options": "defaults, -f 1",
            "l2": "per-step working set (reads + DP arenas, > 1 GB) exceeds the 126 MB L2",
            "parallelism": "read-sharded x%d, no collective%s" % (world, "" if args.workload == "r2c2" else "; %d units of equal predicted work per rank, dealt in snake order" % UNITS_PER_RANK), "lanes_per_gpu": max(1, args.lanes), "e2e_chunk_reads": args.chunk}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(p.get("sm_max_mhz", 1965.0))
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


UNITS_PER_RANK = 4


def rank_units(workload, n_per_gpu, rank, world):
    """This rank's share of the step's batch of world x n_per_gpu reads: a list of units (unit id, names, seqs, first_index).
    r2c2 (uniform reads): one unit, reads [rank n, (rank + 1) n) of the generator.  mixed / mixed_sorted: the batch (sorted by
    nominal length for the latter) is cut into world x UNITS_PER_RANK contiguous units of equal predicted work that are dealt
    in snake order (tidehunter_b200.shard.cut_units / unit_owner), as the sharded front end does."""
    from tidehunter_b200 import synth
    from tidehunter_b200.shard import cut_units, unit_owner, predicted_work
    if workload == "r2c2":
        names, seqs = synth.gen_reads("r2c2", n_per_gpu, start=rank * n_per_gpu)
        return [(rank, names, seqs, rank * n_per_gpu)]
    import numpy as np
    tot = n_per_gpu * world
    nominal = synth.nominal_lengths("mixed", tot)
    order = np.argsort(nominal, kind="stable") if workload == "mixed_sorted" else np.arange(tot)
    units = cut_units(predicted_work(nominal[order]), world * UNITS_PER_RANK if world > 1 else 1)   # one process: one th_host_run over everything
    out = []
    for u, (lo, hi) in enumerate(units):
        if unit_owner(u, world) == rank:
            names, seqs = synth.gen_reads_at("mixed", order[lo:hi])
            out.append((u, names, seqs, lo))
    return out


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation on the host cores
# ---------------------------------------------------------------------------------------------------
def _ref_runner():
    """Returns (kind, fn(names, seqs, threads) -> seconds).  'reference' = unmodified TideHunter binary
    built by oracle/Makefile into oracle/_ref/; 'port' = the oracle's C restatement (only if that binary
    is missing)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    if os.path.exists(O.REF_BIN):
        def run(names, seqs, threads, keep=None):
            with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
                path = os.path.join(td, "sample.fa")
                O.write_fasta(path, names, seqs)
                opath = os.path.join(td, "out.fa") if keep is not None else os.devnull
                t0 = time.perf_counter()
                subprocess.run([O.REF_BIN, "-t", str(threads), "-f", "1", "-o", opath, path], stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL, check=True)
                dt = time.perf_counter() - t0
                if keep is not None:
                    with open(opath, "rb") as f:
                        keep.append(f.read())
                return dt
        return "reference", run

    def run_port(names, seqs, threads, keep=None):
        t0 = time.perf_counter()
        out = O.run_batch(names, seqs, O.default_para(out_fmt=1), threads=threads)[0]
        dt = time.perf_counter() - t0
        if keep is not None:
            keep.append(out)
        return dt
    return "port", run_port


def cpu_sample_size(run, threads, target_s):
    """Calibrate on a few reads, then size the sample for ~target_s seconds of wall time."""
    from tidehunter_b200 import synth
    n0 = max(4 * threads, 32)
    names, seqs = synth.gen_reads("r2c2", n0, start=900000)
    rate = n0 / max(run(names, seqs, threads), 1e-3)
    n1 = int(min(max(rate * 1.5, n0), 4096))          # second pass long enough to amortise start-up
    names, seqs = synth.gen_reads("r2c2", n1, start=900000)
    rate = n1 / max(run(names, seqs, threads), 1e-3)
    return int(min(max(rate * target_s, 2 * threads), 16384))


def reference_arm(args):
    """bench.py --impl reference: the reference's own CPU implementation, all host threads, same
    workload / metric / unit; each step is a bounded sample of the batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from tidehunter_b200 import synth
    cores = os.cpu_count() or 1
    kind, run = _ref_runner()
    n = cpu_sample_size(run, cores, 4.0)
    if args.workload == "r2c2":
        names, seqs = synth.gen_reads("r2c2", n, start=0)
    else:
        units = rank_units(args.workload, args.reads, 0, max(1, args.gpus))   # rank 0's units: a cross-section of the batch
        names = sum((u[1] for u in units), [])[:n]; seqs = sum((u[2] for u in units), [])[:n]
        n = len(seqs)
    bases = synth.total_bases(seqs)
    for _ in range(args.warmup):
        run(names[: max(n // 4, 1)], seqs[: max(n // 4, 1)], cores)
    times = [run(names, seqs, cores) for _ in range(args.steps)]
    tot = sum(times)
    value = n * args.steps / tot
    line = {
        "impl": "reference", "metric": "reads/s", "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16/int32", "data": "synthetic",
        "config": make_config(args, max(1, args.gpus)),
        "sample_reads_per_step": n, "sample_bases_per_step": bases,
        "gbp_per_s": bases * args.steps / tot / 1e9,
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": kind,
                         "sample": "first %d reads (%d bases) of the workload per step, TideHunter -t %d -f 1" % (n, bases, cores)},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def config_legs(device):
    """Short end-to-end legs over the other BASELINE.json config shapes (configs[2..4]): host buffers in, records out
    through th_host_run; next to the unmodified reference (all host cores) on a sample of the same reads, with the outputs
    of that sample compared byte for byte.  Rank 0 at N = 1 only; a few seconds per shape."""
    import gzip
    import hashlib
    import tidehunter_b200 as T
    from tidehunter_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    if not os.path.exists(O.REF_BIN):
        return {"unavailable": "oracle/_ref/TideHunter is not built"}
    with gzip.open(os.path.join(ROOT, "tests", "golden", "golden.json.gz"), "rt") as f:
        ad = json.load(f)["adapters"]
    five, three = ad["five"], ad["three"]
    cores = os.cpu_count() or 1
    legs = [
        ("configs[2] short units 50-200 bp x 20-50 copies, 10% error, -f 2", lambda n: synth.gen_reads("short", n, start=700000), 16384, 1024, ["-f", "2"], dict(out_fmt=2)),
        ("configs[3] long units 4-5 kb x 2-4 copies, 20% error, -f 2", lambda n: synth.gen_reads("long", n, start=700000), 8192, 320, ["-f", "2"], dict(out_fmt=2)),
        ("configs[4] adapters -5 -3 -u -f 2 (unit output)", lambda n: synth.gen_reads("r2c2", n, start=700000, adapters=(five, three)), 8192, 768, ["ADAPTERS", "-u", "-f", "2"],
         dict(out_fmt=2, five_seq=five, three_seq=three, only_unit=1)),
        ("configs[4] adapters -5 -3 -F -f 2 (full-length consensus)", lambda n: synth.gen_reads("r2c2", n, start=700000, adapters=(five, three), three_rc=True), 8192, 512,
         ["ADAPTERS", "-F", "-f", "2"], dict(out_fmt=2, five_seq=five, three_seq=three, only_full_length=1)),
    ]
    rows = []
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        p5, p3 = os.path.join(td, "5.fa"), os.path.join(td, "3.fa")
        with open(p5, "w") as f:
            f.write(">5\n%s\n" % five)
        with open(p3, "w") as f:
            f.write(">3\n%s\n" % three)
        for tag, gen, n, ns, argv, kw in legs:
            names, seqs = gen(n)
            bases = synth.total_bases(seqs)
            th = T.TideHunter(device=device, **kw)
            th.run(names, seqs)                         # warm-up (buffers, slabs)
            ts = []
            for _ in range(2):
                t0 = time.perf_counter()
                out = th.run(names, seqs)
                ts.append(time.perf_counter() - t0)
            ours_s = th.run(names[:ns], seqs[:ns])
            failed = th.failed_tasks()
            th.close()
            dt = min(ts)
            path = os.path.join(td, "s.fa")
            O.write_fasta(path, names[:ns], seqs[:ns])
            argv2 = sum((["-5", p5, "-3", p3] if a == "ADAPTERS" else [a] for a in argv), [])
            t0 = time.perf_counter()
            ref = subprocess.run([O.REF_BIN, "-t", str(cores)] + argv2 + [path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            t_ref = time.perf_counter() - t0
            rows.append({"config": tag, "reads_per_step": n, "bases_per_step": bases, "e2e": {"value": round(n / dt, 1), "unit": "reads/s", "gbp_per_s": round(bases / dt / 1e9, 4), "ms_per_step": round(1e3 * dt, 1)},
                         "output_bytes": len(out), "failed_tasks": failed,
                         "cpu_baseline": {"value": round(ns / t_ref, 1), "unit": "reads/s", "cores": cores, "kind": "reference", "sample": "first %d reads, TideHunter -t %d %s" % (ns, cores, " ".join(argv).replace("ADAPTERS", "-5 .. -3 .."))},
                         "parity": {"reads": ns, "identical": bool(ours_s == ref), "md5_reference": hashlib.md5(ref).hexdigest()}})
    return rows


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(5)
            except Exception:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw)}


STAGES = ("pack", "seed", "chain", "select", "partition", "poa", "ksw", "d2h")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=32768, help="reads per GPU per step")
    ap.add_argument("--lanes", type=int, default=4, help="GPU contexts (streams) the step's reads are dealt to")
    ap.add_argument("--chunk", type=int, default=4096, help="reads per chunk of the end-to-end leg (th_host_run)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="r2c2", choices=sorted(WORKLOADS), help="r2c2 = BASELINE configs[1] (the contract's line); mixed* = mixed read lengths, dealt by bases")
    ap.add_argument("--no-configs", action="store_true", help="skip the short legs over BASELINE configs[2..4]")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly one line, the JSON: whatever libraries print to file descriptor 1 (NCCL's version line) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import tidehunter_b200 as T
    from tidehunter_b200 import synth
    from tidehunter_b200.shard import ordered_gather
    if not torch.cuda.is_available() or T.gpu_lib().th_gpu_device_count() <= 0:
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    gloo = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        gloo = dist.new_group(backend="gloo")  # host-side ordered gather only; no data-path collective

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=gloo)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # this rank's batch: read indices [rank*reads, (rank+1)*reads) of the seeded generator
    n = args.reads
    L = max(1, args.lanes)
    units = rank_units(args.workload, n, rank, world)
    names = sum((u[1] for u in units), []); seqs = sum((u[2] for u in units), [])
    n = len(seqs)                                      # mixed workloads: equal predicted work per rank, not equal read counts
    bases = synth.total_bases(seqs)

    # ---------------- device-resident leg (value) ----------------
    # The batch is dealt to L contexts ("lanes": own stream, own buffers) in contiguous blocks; one host thread per
    # lane drives its context, so the lanes' kernels overlap on the GPU exactly as they do under th_host_run.
    import threading
    per_lane = (n + L - 1) // L
    ctxs = [T.GpuContext(device=local) for _ in range(L)]
    for k, c in enumerate(ctxs):
        c.upload(seqs[k * per_lane:(k + 1) * per_lane])

    def run_lanes(steps, collect=None):
        def work(k):
            ctxs[k].mark(0)                       # CUDA event on the lane's stream: start of its first step
            for _ in range(steps):
                r = ctxs[k].process_resident()
                if collect is not None:
                    collect[k].append(r.stats.as_dict())
            ctxs[k].mark(1)                       # ... and the end of its last step
        th = [threading.Thread(target=work, args=(k,)) for k in range(L)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    run_lanes(args.warmup)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    lane_stats = [[] for _ in range(L)]
    run_lanes(args.steps, lane_stats)
    barrier()
    dt_host = time.perf_counter() - t0
    clocks = sampler.stop()
    # device time of the timed region: latest end mark minus earliest start mark over the lanes' streams (CUDA events)
    dt = max(ctxs[a].elapsed_ms(0, ctxs[b], 1) for a in range(L) for b in range(L)) * 1e-3
    dt = max_over_ranks(dt)
    dt_host = max_over_ranks(dt_host)
    acc = {k: 0.0 for k in STAGES + ("total",)}
    cnt = {}
    launches = 0
    for k in range(L):
        for s in lane_stats[k]:
            for st in STAGES + ("total",):
                acc[st] += s["ms_" + st]
            launches += s["n_launches"]
    for key in ("n_bases", "n_hits", "n_chain_evals", "n_poa_cells", "n_poa_rows", "n_ksw_cells", "n_ksw_cells_full", "n_tasks"):
        cnt[key] = sum(lane_stats[k][-1][key] for k in range(L))   # per step, all lanes
    n_tasks = cnt["n_tasks"]
    # one lane alone (serial stages, nothing co-running): the per-kernel times comparable with the ncu launch list
    serial = {k: 0.0 for k in STAGES}
    serial_range = {}
    serial_cnt = {}
    if L > 1:
        runs = [ctxs[0].process_resident().stats.as_dict() for _ in range(5)]
        for st in STAGES:   # median of five launches: the per-launch time of the POA kernel varies between launches (63 - 72 ms, once 122 ms, on the same batch)
            serial[st] = statistics.median(r["ms_" + st] for r in runs)
            serial_range[st] = [round(min(r["ms_" + st] for r in runs), 3), round(max(r["ms_" + st] for r in runs), 3)]
        serial_cnt = runs[-1]
    for c in ctxs:
        c.close()

    # ---------------- end-to-end leg through the host layer (e2e) ----------------
    # The reads are handed over as the C entry point takes them (arrays of pointers and lengths into host memory, built
    # once: tidehunter_b200.Batch); every step then runs th_host_run on those HOST buffers -- staging into pinned memory,
    # H2D, all kernels, D2H, record formatting -- and reads the output text where the library leaves it.
    th = T.TideHunter(device=local, out_fmt=1, chunk_reads=args.chunk, lanes=L)
    from tidehunter_b200.shard import ordered_gather_units
    batches = [(u[0], T.Batch(u[1], u[2]), u[3]) for u in units]
    single = len(batches) == 1

    def run_units():
        if single:   # the library's own output buffer, no copy
            return [th.run(batches[0][1], first_index=batches[0][2], copy=False)]
        return [th.run(b, first_index=lo) for _, b, lo in batches]   # one th_host_run per unit, in input order
    for _ in range(2):
        run_units()
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    out_bytes = 0
    t_run = t_gather = 0.0
    for _ in range(args.steps):
        ta = time.perf_counter()
        texts = run_units()
        tb = time.perf_counter()
        parts = ordered_gather(texts[0], rank, world, gloo) if single else ordered_gather_units(texts, [b[0] for b in batches], rank, world, gloo)
        if parts is not None:
            out_bytes = sum(len(p) for p in parts)
        del parts
        t_gather += time.perf_counter() - tb; t_run += tb - ta
        s = th.stats()
        h2d += s["h2d_bytes"]; d2h += s["d2h_bytes"]
    barrier()
    dt_e2e = max_over_ranks(time.perf_counter() - t0)
    t_run = max_over_ranks(t_run); t_gather = max_over_ranks(t_gather)

    tot_reads = sum_over_ranks(n) * args.steps
    tot_bases = sum_over_ranks(bases) * args.steps
    value = tot_reads / dt
    e2e = tot_reads / dt_e2e

    # ---------------- roofline for the dominant kernel + per-kernel tables ----------------
    # `kernels`: one lane alone (each kernel has the whole GPU; comparable with the ncu launch list in profiles/).
    # `kernels_overlapped`: the same CUDA-event brackets inside the timed region, where the lanes' kernels share the SMs;
    # times are per launch (one launch of each kernel per lane per step), so they include the co-running slowdown.
    hbm_peak, peak_src, sm_max = load_peaks()
    unit_name = {"pack": "bases", "seed": "hits", "chain": "pair_evals", "poa": "cells", "ksw": "cells"}
    key_of = {"pack": "n_bases", "seed": "n_hits", "chain": "n_chain_evals", "poa": "n_poa_cells", "ksw": "n_ksw_cells"}

    def table(ms_per_launch, counts):
        out = {}
        tot = sum(ms_per_launch.values())
        for k in STAGES:
            e = {"ms_per_launch": round(ms_per_launch[k], 3), "share": round(ms_per_launch[k] / tot, 4) if tot > 0 else None}
            if k in key_of and ms_per_launch[k] > 0:
                e["unit"] = unit_name[k]
                e["g_units_per_s"] = round(counts[key_of[k]] / (ms_per_launch[k] * 1e-3) / 1e9, 3)
                if k == "ksw" and counts.get("n_ksw_cells_full"):
                    # `cells` = cells computed (certified bands of the identity alignments, cut extension matrices); the
                    # reference computes the full matrices: their cells per second is the rate comparable with a CPU's GCUPS
                    e["computed_share_of_full_matrices"] = round(counts["n_ksw_cells"] / counts["n_ksw_cells_full"], 4)
                    e["g_full_matrix_cells_per_s"] = round(counts["n_ksw_cells_full"] / (ms_per_launch[k] * 1e-3) / 1e9, 3)
            out[k] = e
        return out

    over_ms = {k: acc[k] / (args.steps * L) for k in STAGES}
    over_cnt = {key: cnt[key] / L for key in list(key_of.values()) + ["n_ksw_cells_full"]}
    kernels_over = table(over_ms, over_cnt)
    if L > 1:
        kernels = table(serial, serial_cnt)
        for k, mm in serial_range.items():   # the median is what the fractions use; the spread of the five launches is shown beside it
            kernels[k]["ms_per_launch_min_max"] = mm
        ser_ms, ser_cnt = serial, serial_cnt
    else:
        kernels, ser_ms, ser_cnt = kernels_over, over_ms, over_cnt
    dom = max(("poa", "ksw", "chain", "seed", "pack"), key=lambda k: ser_ms[k])
    # Algorithmic HBM bytes per unit: POA stores 7 B per banded cell (H, E1, E2 as int16 + one code
    # byte) and reads the three planes once more as a predecessor row or in the backtrack (13 B/cell); ksw keeps its rows in
    # registers (boundary hand-off only: 16 B per target row per 256-column block, ~0.06 B/cell); chaining reads 12 B per
    # evaluated predecessor (L1/L2 hits); seeding reads L/4 + L/8 bytes and writes 8 B per hit; packing 1 B in, 1.375 B out.
    bytes_per_unit = {"poa": 13.0, "ksw": 16.0 / 256, "chain": 12.0, "seed": None, "pack": 1.0 + 1.0 + 0.25 + 0.125}
    bound_of = {"pack": "hbm", "seed": "hbm", "chain": "int", "poa": "int", "ksw": "int"}
    ncu = load_ncu_constants()
    n_sm = 148
    # integer roof: the ALU pipe issues 2 warp-instructions per clock per SM for the packed 16x2 DP operations
    # (VIADD.16x2, VIMNMX.S16x2, VIMNMX3, PRMT, IMAD: tools/microbench/alu_peak.cu, profiles/r1_alu_peak.json)
    int_peak = n_sm * sm_max * 1e6 * 2.0          # ALU-pipe warp-instructions per second
    issue_peak = n_sm * sm_max * 1e6 * 4.0        # issue slots per second

    def alg_bytes_k(k, counts):
        if k == "seed":
            return counts["n_bases"] * (0.25 + 0.125) + 8.0 * counts["n_hits"]
        return counts[key_of[k]] * bytes_per_unit[k]

    def annotate(tab, ms, counts):
        """hbm_frac for every kernel; int_frac / issue_frac where the ncu constants know the instruction mix."""
        for k in key_of:
            if ms[k] <= 0:
                continue
            e = tab[k]
            e["bound"] = bound_of[k]
            e["hbm_frac"] = round(alg_bytes_k(k, counts) / (ms[k] * 1e-3) / 1e9 / hbm_peak, 5)
            c = ncu.get(k)
            if c and bound_of[k] == "int":
                units = counts[key_of[k]]
                e["int_frac"] = round(c["alu_inst_per_unit"] * units / (ms[k] * 1e-3) / int_peak, 4)
                e["issue_frac"] = round(c["warp_inst_per_unit"] * units / (ms[k] * 1e-3) / issue_peak, 4)
                e["warp_inst_per_unit"] = c["warp_inst_per_unit"]

    annotate(kernels, ser_ms, ser_cnt)
    annotate(kernels_over, over_ms, over_cnt)
    alg_bytes = alg_bytes_k(dom, ser_cnt)
    achieved = alg_bytes / (ser_ms[dom] * 1e-3) / 1e9 if ser_ms[dom] > 0 else 0.0
    achieved_over = alg_bytes_k(dom, over_cnt) / (over_ms[dom] * 1e-3) / 1e9 if over_ms[dom] > 0 else 0.0
    cdom = ncu.get(dom, {})
    traffic = cdom.get("dram_bytes_per_unit")
    roofline = {"kernel": {"poa": "poa_kernel", "ksw": "ksw_pair_kernel+ksw_ext_kernel", "chain": "chain_dp_kernel", "seed": "seed_kernel", "pack": "pack_kernel"}[dom],
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": (traffic * ser_cnt[key_of[dom]] / 1e9) if traffic else None, "traffic_unit": "GB per launch (ncu dram bytes per unit x units of this launch)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ser_ms[dom],
                "measured": "CUDA events on the library's stream around the kernel, one lane alone (after the timed region)" if L > 1 else "CUDA events on the library's stream inside the timed region",
                "in_timed_region": {"ms_per_launch": over_ms[dom], "achieved": achieved_over, "frac": achieved_over / hbm_peak, "lanes_co_running": L},
                "binding": bound_of[dom],
                "int_frac": kernels[dom].get("int_frac"), "issue_frac": kernels[dom].get("issue_frac"),
                "int_peak": {"value": int_peak / 1e9, "unit": "G ALU-pipe warp-instructions/s (2 per clock per SM, profiles/r1_alu_peak.json)"},
                "ncu_constants": "profiles/r2_ncu_constants.json" if ncu else None,
                "note": "integer DP kernel: `frac` is the HBM fraction the contract asks for and shows the kernel is not bandwidth-bound; "
                        "`int_frac` = ALU-pipe warp-instructions per unit (ncu) x units / live kernel time / the measured 2-per-clock pipe rate is the roof that "
                        "applies; `kernels` carries both fractions for every kernel"}

    line = {
        "metric": "reads/s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "host_clock_ms_per_step": 1e3 * dt_host / args.steps,
        "timing": "CUDA events on the lanes' own streams (latest end - earliest start), max over ranks; host clock between synchronised barriers alongside",
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16/int32", "data": "synthetic",
        "config": make_config(args, world),
        "reads_this_rank_per_step": n, "bases_this_rank_per_step": bases,
        "gbp_per_s": tot_bases / dt / 1e9,
        "poa_gcups": kernels["poa"].get("g_units_per_s"), "ksw_gcups": kernels["ksw"].get("g_units_per_s"),
        "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                "ms_per_step": 1e3 * dt_e2e / args.steps, "output_bytes_per_step": out_bytes, "gbp_per_s": tot_bases / dt_e2e / 1e9,
                "th_host_run_ms_per_step_max_rank": round(1e3 * t_run / args.steps, 1), "ordered_gather_ms_per_step_max_rank": round(1e3 * t_gather / args.steps, 1)},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels, "kernels_overlapped": kernels_over,
        "lane_ms_per_step": round(acc["total"] / (args.steps * L), 3), "poa_tasks_per_step": n_tasks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind, run = _ref_runner()
        cores = os.cpu_count() or 1
        ns = cpu_sample_size(run, cores, 15.0)
        ns = min(ns, n)
        ref_out = []
        t = run(names[:ns], seqs[:ns], cores, keep=ref_out)
        # parity at bench scale: the reference's output for the sample against ours for the same reads (untimed)
        import hashlib
        ours = th.run(names[:ns], seqs[:ns])
        line["parity"] = {"reads": ns, "identical": bool(ours == ref_out[0]), "against": kind, "records": ours.count(b">"),
                          "md5_ours": hashlib.md5(ours).hexdigest(), "md5_reference": hashlib.md5(ref_out[0]).hexdigest()}
        line["cpu_baseline"] = {"value": ns / t, "unit": "reads/s", "cores": cores, "kind": kind,
                                "sample": "first %d reads of the step's batch (%d bases), TideHunter -t %d -f 1, %.1f s" % (ns, synth.total_bases(seqs[:ns]), cores, t)}
    th.close()
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.no_configs and args.workload == "r2c2":
        line["configs"] = config_legs(local)
    if rank == 0:
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier(group=gloo)
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
