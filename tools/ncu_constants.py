#!/usr/bin/env python
"""profiles/r2_ncu_constants.json from ncu --set full captures (tools/gpu_job.sh ncu ...): per stage, warp-instructions,
ALU-pipe warp-instructions and DRAM bytes per algorithmic unit, summed over the stage's kernels.  bench.py multiplies them
with the units of its own launches and divides by its own CUDA-event time.
usage: python tools/ncu_constants.py stage=report.ncu-rep[,report2...]:stepstats.log ...   (stage in poa, ksw, chain)
The log is the capture's own stdout (tools/profile_step.py prints the step's work counters); ALU-pipe instructions =
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active x 2 per clock x sm__cycles_active.avg x SMs."""
import ast
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT_KEY = {"poa": "n_poa_cells", "ksw": "n_ksw_cells", "chain": "n_chain_evals"}
N_SM = 148


def launches(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))

        def val(k, scale={"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}):
            return float(d[k].replace(",", "")) * scale.get(u[k], 1.0)
        res.append({"kernel": d["Kernel Name"].split("(")[0], "ms": val("gpu__time_duration.sum"), "inst": val("smsp__inst_executed.sum"),
                    "alu_inst": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") / 100.0 * 2.0 * val("sm__cycles_active.avg") * N_SM,
                    "dram": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                    "alu_pct": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), "ipc": val("sm__inst_executed.avg.per_cycle_elapsed"),
                    "regs": val("launch__registers_per_thread"), "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active")})
    return res


def step_stats(log):
    for ln in open(log, errors="replace"):
        ln = ln.strip()
        if ln.startswith("{'ms_h2d'"):
            return ast.literal_eval(ln)
    raise SystemExit("no step statistics in " + log)


def main():
    path = os.path.join(ROOT, "profiles", "r2_ncu_constants.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for a in sys.argv[1:]:
        stage, rest = a.split("=", 1)
        reps, log = rest.rsplit(":", 1)
        ls = sum((launches(r) for r in reps.split(",")), [])
        units = step_stats(log)[UNIT_KEY[stage]]
        out[stage] = {"unit": UNIT_KEY[stage], "units_in_capture": units,
                      "warp_inst_per_unit": round(sum(l["inst"] for l in ls) / units, 5), "alu_inst_per_unit": round(sum(l["alu_inst"] for l in ls) / units, 5),
                      "dram_bytes_per_unit": round(sum(l["dram"] for l in ls) / units, 4), "kernels": ls,
                      "source": [os.path.basename(r) for r in reps.split(",")]}
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps({k: {x: v[x] for x in ("warp_inst_per_unit", "alu_inst_per_unit", "dram_bytes_per_unit")} for k, v in out.items()}))


if __name__ == "__main__":
    main()
