mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final4.json 2> gpurun_out/bench_final4.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final4.json 2> gpurun_out/bench_ref_final4.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_final4.json"))
print(round(d["value"]), round(d["e2e"]["value"]), d["parity"]["identical"], round(d["cpu_baseline"]["value"]), d["roofline"]["frac"], {k:v["ms_per_launch"] for k,v in d["kernels"].items()})
r=json.load(open("gpurun_out/bench_ref_final4.json")); print(round(r["value"]))
PY
