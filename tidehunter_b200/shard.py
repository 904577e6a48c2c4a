"""Read-batch sharding across ranks and the host-side ordered gather.

Reads are independent (the reference's only parallelism is a thread pool over reads, src/main.c:273-291),
so rank r of W processes the contiguous block shard_range(n, r, W) of every batch on its own GPU and
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


_gather_seq = 0


def _same_node(world):
    """All ranks on one node (the only layout the reference has: one process, one box)?"""
    lw = os.environ.get("LOCAL_WORLD_SIZE")
    return lw is not None and int(lw) == world and os.path.isdir("/dev/shm") and os.environ.get("TH_GATHER", "shm") == "shm"


def ordered_gather(payload, rank, world, group=None, dst=0):
    """Gather one bytes payload per rank on `dst`, in rank order.  Returns the list on dst, None elsewhere.

    On one node the texts travel through tmpfs files (/dev/shm) bracketed by two host-side barriers: memcpy speed,
    no pickling of tens of MB per rank through the loopback.  Otherwise torch.distributed.gather_object over the
    gloo group.  Either way it is a host-side gather; nothing touches the GPUs or NCCL."""
    global _gather_seq
    if world == 1:
        return [payload]
    import torch.distributed as dist
    if not _same_node(world):
        out = [None] * world if rank == dst else None
        dist.gather_object(payload, out, dst=dst, group=group)
        return out
    _gather_seq += 1
    tag = "th_b200_%s_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "run"), _gather_seq)
    path = os.path.join("/dev/shm", "%s_r%d" % (tag, rank))
    if rank != dst:
        with open(path, "wb") as f:
            f.write(payload)
    dist.barrier(group=group)            # every rank's file is complete
    out = None
    if rank == dst:
        out = []
        for r in range(world):
            if r == dst:
                out.append(payload)
                continue
            pr = os.path.join("/dev/shm", "%s_r%d" % (tag, r))
            with open(pr, "rb") as f:
                out.append(f.read())
            os.unlink(pr)
    return out


def run_sharded(th, names, seqs, rank, world, group=None):
    """Process this rank's block of (names, seqs) with `th` (a tidehunter_b200.TideHunter) and gather the
    text on rank 0 in input order.  Returns bytes on rank 0, None elsewhere."""
    lo, hi = shard_range(len(seqs), rank, world)
    text = th.run(names[lo:hi], seqs[lo:hi])
    parts = ordered_gather(text, rank, world, group)
    return b"".join(parts) if parts is not None else None
