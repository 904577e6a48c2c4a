"""Host-side multi-GPU logic on CPU: block sharding tiles the batch in order, and the gloo ordered gather
(world_size 2) reproduces the single-process output.  The per-rank "engine" here is the CPU oracle,
standing in for the GPU context only to exercise the sharding / gather code."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_tile_in_order():
    from tidehunter_b200.shard import shard_range
    for n in (0, 1, 7, 8, 4096, 100003):
        for w in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_work_balanced_ranges_tile_in_order_and_balance_bases():
    """shard_range_by_work: contiguous, in order, and no rank's bases exceed the ideal share by more than one read --
    also for length-sorted input, where equal read counts would give the last rank several times the work."""
    import numpy as np
    from tidehunter_b200.shard import shard_range, shard_range_by_work
    rng = np.random.default_rng(7)
    for n in (0, 1, 3, 100, 5000):
        for w in (1, 2, 3, 8):
            for order in ("random", "sorted"):
                lens = rng.integers(1800, 23700, n)
                if order == "sorted":
                    lens = np.sort(lens)
                blocks = [shard_range_by_work(lens, r, w) for r in range(w)]
                assert blocks[0][0] == 0 and blocks[-1][1] == n
                assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
                if n >= 100:
                    share = [int(lens[a:b].sum()) for a, b in blocks]
                    assert max(share) <= lens.sum() / w + lens.max()
    lens = np.sort(rng.integers(1800, 23700, 5000))
    by_count = max(int(lens[a:b].sum()) for a, b in (shard_range(5000, r, 8) for r in range(8)))
    by_work = max(int(lens[a:b].sum()) for a, b in (shard_range_by_work(lens, r, 8) for r in range(8)))
    assert by_count > 1.5 * by_work


def test_snake_dealt_units_balance_a_sorted_batch():
    """cut_units + unit_owner: the units tile the input in order, every rank gets units_per_rank of them, and on a
    length-sorted batch whose cost per base grows with the read length the busiest rank is close to the mean -- contiguous
    parts of equal bases leave the last rank with several times the work."""
    import numpy as np
    from tidehunter_b200.shard import cut_units, unit_owner, predicted_work, shard_range_by_work
    rng = np.random.default_rng(11)
    lens = np.sort(rng.choice([1813, 2104, 2509, 3026, 3254, 3805, 5211, 6463, 9390, 23611], 40000))
    cost = lens.astype(np.float64) * np.maximum(lens - 2500, 0)       # nothing below ~2 units, superlinear above
    world, upr = 8, 4
    units = cut_units(predicted_work(lens), world * upr)
    assert units[0][0] == 0 and units[-1][1] == len(lens) and all(units[i][1] == units[i + 1][0] for i in range(len(units) - 1))
    owners = [unit_owner(u, world) for u in range(len(units))]
    assert owners[:world] == list(range(world)) and owners[world:2 * world] == list(range(world - 1, -1, -1))
    assert all(owners.count(r) == upr for r in range(world))
    load = np.zeros(world)
    for u, (lo, hi) in enumerate(units):
        load[owners[u]] += cost[lo:hi].sum()
    contiguous = np.array([cost[slice(*shard_range_by_work(predicted_work(lens), r, world))].sum() for r in range(world)])
    assert load.max() < 0.6 * contiguous.max()
    assert load.max() / load.mean() < 1.6 < contiguous.max() / contiguous.mean()


def test_unit_gather_orders_by_unit_id():
    from tidehunter_b200.shard import ordered_gather_units
    out = ordered_gather_units([b"bb", b"", b"dddd"], [1, 2, 3], 0, 1)
    assert [bytes(x) for x in out] == [b"bb", b"", b"dddd"]


def _worker(rank, world, port, q, shm=False):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    if shm:  # what torchrun exports on a single node: the gather then goes through /dev/shm files
        os.environ["LOCAL_WORLD_SIZE"] = str(world); os.environ["MASTER_PORT"] = str(port)
    else:
        os.environ.pop("LOCAL_WORLD_SIZE", None)
    import torch.distributed as dist
    import oracle_py as O
    from tidehunter_b200 import synth
    from tidehunter_b200.shard import run_sharded
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)

    class Engine:  # same .run(names, seqs) -> bytes interface as tidehunter_b200.TideHunter
        def run(self, names, seqs, first_index=None):
            return O.run_batch(names, seqs, O.default_para(out_fmt=2))[0]
    names, seqs = synth.gen_reads("short", 9)
    out = run_sharded(Engine(), names, seqs, rank, world)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shm", [False, True])
def test_two_rank_ordered_gather_matches_single_process(oracle, shm):
    import torch.multiprocessing as mp
    from tidehunter_b200 import synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q, shm)) for r in range(2)]
    for p in ps:
        p.start()
    out = q.get(timeout=120)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    names, seqs = synth.gen_reads("short", 9)
    assert out == oracle.run_batch(names, seqs, oracle.default_para(out_fmt=2))[0]
    assert out.count(b"\n") >= 9
