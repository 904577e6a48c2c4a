#!/usr/bin/env python
"""Multi-GPU parity: W ranks (one per GPU, torchrun) each process their units of one seeded batch (tidehunter_b200.shard:
units of equal predicted work dealt in snake order) through the host layer over the C ABI; rank 0 gathers the texts in
unit order (gloo, host side) and compares the result byte for byte with the unmodified reference run on the whole batch.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P tools/sharded_parity.py [n_reads] [shape]

Test infrastructure (executes oracle/_ref on rank 0); nothing here is on the product path.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6144
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist
    import tidehunter_b200 as T
    from tidehunter_b200 import synth
    from tidehunter_b200.shard import run_sharded
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="gloo")
        group = dist.group.WORLD
    shape = sys.argv[2] if len(sys.argv) > 2 else "mixed"   # mixed read lengths: the units of run_sharded differ in size
    names, seqs = synth.gen_reads(shape, n, start=700000)
    th = T.TideHunter(device=local, out_fmt=2)
    text = run_sharded(th, names, seqs, rank, world, group)
    th.close()
    rc = 0
    if rank == 0:
        import oracle_py as O
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
            path = os.path.join(td, "in.fa")
            O.write_fasta(path, names, seqs)
            ref = subprocess.run([O.REF_BIN, "-t", str(os.cpu_count() or 1), "-f", "2", path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        rep = {"world_size": world, "reads": n, "identical": text == ref, "bytes": len(ref), "md5_reference": hashlib.md5(ref).hexdigest(),
               "md5_ours": hashlib.md5(text).hexdigest(), "options": "-f 2", "shape": shape, "gather": "host-side, unit order (tmpfs files + gloo barrier on one node, gather_object otherwise); no data-path collective"}
        print(json.dumps(rep))
        rc = 0 if text == ref else 1
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
