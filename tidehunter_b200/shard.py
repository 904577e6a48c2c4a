"""Read-batch sharding across ranks and the host-side ordered gather.

Reads are independent (the reference's only parallelism is a thread pool over reads, src/main.c:273-291),
so rank r of W processes a contiguous block of every batch on its own GPU (shard_range_by_work: equal bases, not equal
read counts) and
there is no data-path collective.  Output order = input order: rank 0 concatenates the ranks' output
texts in rank order, exactly what mini_tandem_output (src/main.c:214-271) prints for the whole batch.
The gather runs over a gloo (host) group; NCCL / NVLink are not on the data path.
"""
import os


def shard_range(n, rank, world):
    """Contiguous, balanced block of [0, n) for `rank`; blocks of consecutive ranks tile [0, n) in order."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_range_by_work(work, rank, world):
    """Contiguous block of [0, len(work)) for `rank` such that the blocks of consecutive ranks tile the input in order and
    carry about equal total `work` (per read: its length in bases -- seeding, chaining and consensus all grow with it).
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


def run_sharded(th, names, seqs, rank, world, group=None):
    """Process this rank's block of (names, seqs) with `th` (a tidehunter_b200.TideHunter) and gather the
    text on rank 0 in input order.  Returns bytes on rank 0, None elsewhere."""
    lo, hi = shard_range_by_work([len(x) for x in seqs], rank, world)
    text = th.run(names[lo:hi], seqs[lo:hi], first_index=lo)   # global read index: the FASTQ quality slot follows it
    parts = ordered_gather(text, rank, world, group)
    return b"".join(bytes(p) for p in parts) if parts is not None else None
