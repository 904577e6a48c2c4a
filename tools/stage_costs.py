"""Per-stage device time per base for several read shapes (one lane, device-resident): where a workload's cost per base
comes from.  usage: python tools/stage_costs.py shape:n [shape:n ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tidehunter_b200 as T
from tidehunter_b200 import synth

for a in sys.argv[1:]:
    shape, n = a.split(":")
    names, seqs = synth.gen_reads(shape, int(n))
    ctx = T.GpuContext()
    ctx.upload(seqs)
    for _ in range(2):
        d = ctx.process_resident().stats.as_dict()
    ctx.close()
    b = d["n_bases"]
    print(json.dumps({"shape": shape, "reads": int(n), "Mbases": round(b / 1e6, 1), "ns_per_base": round(d["ms_total"] * 1e6 / b, 3), "tasks": d["n_tasks"],
                      "poa_cells_per_base": round(d["n_poa_cells"] / b, 1), "ksw_cells_per_base": round(d["n_ksw_cells"] / b, 1),
                      "chain_evals_per_base": round(d["n_chain_evals"] / b, 1), "ms": {k[3:]: round(v, 1) for k, v in d.items() if k.startswith("ms_")}}), flush=True)
