"""Read-batch sharding across ranks and the host-side ordered gather.

Reads are independent (the reference's only parallelism is a thread pool over reads, src/main.c:273-291),
so a batch is cut into contiguous units of equal predicted work that are dealt to the ranks in snake order (cut_units,
unit_owner), every rank processes its units on its own GPU, and
there is no data-path collective.  Output order = input order: rank 0 concatenates the units' output
texts in unit order, exactly what mini_tandem_output (src/main.c:214-271) prints for the whole batch.
The gather runs over a gloo (host) group; NCCL / NVLink are not on the data path.
"""
import os


def shard_range(n, rank, world):
    """Contiguous, balanced block of [0, n) for `rank`; blocks of consecutive ranks tile [0, n) in order."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


READ_OVERHEAD_BASES = 3000


def predicted_work(lengths):
    """Work of a read, in bases: its length plus a per-read cost.  Calibrated on one B200 from two workloads of different mean
    read length (32,768 reads of 10.1 kb: 534 ms per step; 65,536 reads of 4.1 kb mean: 588 ms): time = 1.2 ns per base
    + 4.0 us per read, i.e. a read costs as much as ~3,300 bases (block-per-read and thread-per-read kernels, launch tails)."""
    import numpy as np
    return np.asarray(lengths, dtype=np.int64) + READ_OVERHEAD_BASES


def shard_range_by_work(work, rank, world):
    """Contiguous block of [0, len(work)) for `rank` such that the blocks of consecutive ranks tile the input in order and
    carry about equal total `work` (per read: predicted_work of its length -- seeding, chaining and consensus all grow with it).
    The reference balances reads over its threads dynamically (src/main.c:273-291); with mixed read lengths equal read
    counts would be unequal work.  Block r ends at the first read where the running total reaches (r + 1) / world of the
    sum, so every rank derives the same cuts from the lengths alone."""
    import numpy as np
    w = np.asarray(work, dtype=np.int64)
    n = len(w)
    if n == 0:
        return 0, 0
    cum = np.cumsum(np.maximum(w, 1))
    tot = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(max(cuts[-1], int(np.searchsorted(cum, -(-tot * r // world), side="left")) + 1))
    cuts.append(n)
    cuts = [min(c, n) for c in cuts]
    return cuts[rank], cuts[rank + 1]


def cut_units(work, n_units):
    """Cuts [0, len(work)) into at most n_units contiguous units of about equal total work; returns [(lo, hi), ...]."""
    n_units = max(1, int(n_units))
    cuts = [shard_range_by_work(work, u, n_units) for u in range(n_units)]
    return [(lo, hi) for lo, hi in cuts if hi > lo]


def unit_owner(u, world):
    """Rank that processes unit u: units are dealt in snake order (0 1 .. W-1, W-1 .. 1 0, 0 1 ..), so that a cost gradient
    along the input -- reads sorted by length, where equal predicted work is not equal time because the cost per base grows
    with the number of copies -- ends up spread over all ranks instead of on the last one ("length-sorted
    bins dealt round-robin"; the reference's threads pull reads dynamically, src/main.c:273-291)."""
    k, r = divmod(u, world)
    return r if k % 2 == 0 else world - 1 - r


_gather_seq = 0


def _same_node(world):
    """All ranks on one node (the only layout the reference has: one process, one box)?"""
    lw = os.environ.get("LOCAL_WORLD_SIZE")
    return lw is not None and int(lw) == world and os.path.isdir("/dev/shm") and os.environ.get("TH_GATHER", "shm") == "shm"


def ordered_gather(payload, rank, world, group=None, dst=0):
    """Gather one bytes payload per rank on `dst`, in rank order.  Returns the list on dst, None elsewhere.

    On one node the texts travel through tmpfs files (/dev/shm, created exclusively) followed by a host-side barrier;
    `dst` maps them (the list then holds read-only memoryviews of shared memory, nothing is copied a second time): memcpy speed,
    no pickling of tens of MB per rank through the loopback.  Otherwise torch.distributed.gather_object over the
    gloo group.  Either way it is a host-side gather; nothing touches the GPUs or NCCL."""
    global _gather_seq
    if world == 1:
        return [payload]
    import torch.distributed as dist
    if not _same_node(world):
        out = [None] * world if rank == dst else None
        dist.gather_object(bytes(payload), out, dst=dst, group=group)
        return out
    _gather_seq += 1
    tag = "th_b200_%s_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "run"), _gather_seq)
    path = os.path.join("/dev/shm", "%s_r%d" % (tag, rank))
    if rank != dst:   # never follow or reuse something already sitting at that name
        fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_EXCL | getattr(os, "O_NOFOLLOW", 0), 0o600)
        with os.fdopen(fd, "wb") as f:
            f.write(payload)
    dist.barrier(group=group)            # every rank's file is complete
    out = None
    if rank == dst:   # map the other ranks' files instead of reading them: the pages already are in this node's memory
        import mmap
        out = []
        for r in range(world):
            if r == dst:
                out.append(payload)
                continue
            pr = os.path.join("/dev/shm", "%s_r%d" % (tag, r))
            with open(pr, "rb") as f:
                size = os.fstat(f.fileno()).st_size
                out.append(memoryview(mmap.mmap(f.fileno(), size, access=mmap.ACCESS_READ)) if size else memoryview(b""))
            os.unlink(pr)   # the mapping keeps the pages until the views are dropped
    return out


def ordered_gather_units(texts, unit_ids, rank, world, group=None, dst=0):
    """Gather of per-unit output texts: every rank hands in the texts of the units it processed (`unit_ids`, ascending) and
    `dst` gets all texts ordered by unit id (a list of bytes-like objects), None elsewhere.  One payload per rank travels
    through ordered_gather (the texts back to back behind a small index), so the transport is the same."""
    import json
    import struct
    index = json.dumps([[int(u), len(t)] for u, t in zip(unit_ids, texts)]).encode()
    payload = b"".join([struct.pack("<q", len(index)), index] + [bytes(t) if not isinstance(t, (bytes, bytearray, memoryview)) else t for t in texts])
    parts = ordered_gather(payload, rank, world, group, dst)
    if parts is None:
        return None
    by_unit = {}
    for p in parts:
        mv = memoryview(p)
        (il,) = struct.unpack("<q", bytes(mv[:8]))
        off = 8 + il
        for u, ln in json.loads(bytes(mv[8:8 + il])):
            by_unit[u] = mv[off:off + ln]
            off += ln
    return [by_unit[u] for u in sorted(by_unit)]


def run_sharded(th, names, seqs, rank, world, group=None, units_per_rank=4):
    """Process this rank's share of (names, seqs) with `th` (a tidehunter_b200.TideHunter) and gather the text on rank 0 in
    input order.  The input is cut into world x units_per_rank contiguous units of equal predicted work, dealt in snake
    order (unit_owner); a unit is one th.run call, told the global index of its first read.  Exact for FASTA / tabular
    output; FASTQ (-f 3/4) is exact up to 4,096 reads (host/th_host.h: the reference's never-rewound quality buffers).  Returns bytes on rank 0, None elsewhere."""
    if world == 1:
        return th.run(names, seqs)
    units = cut_units(predicted_work([len(x) for x in seqs]), world * units_per_rank)
    mine = [u for u in range(len(units)) if unit_owner(u, world) == rank]
    texts = [th.run(names[units[u][0]:units[u][1]], seqs[units[u][0]:units[u][1]], first_index=units[u][0]) for u in mine]
    parts = ordered_gather_units(texts, mine, rank, world, group)
    return b"".join(bytes(p) for p in parts) if parts is not None else None
