mkdir -p gpurun_out
timeout 900 python tools/config_bench.py > gpurun_out/config_bench.log 2>&1; echo "config bench exit $?"
tail -n 5 gpurun_out/config_bench.log | cut -c1-600
python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_final3.json"))
print(round(d["value"]), round(d["e2e"]["value"]), d["parity"]["identical"], round(d["cpu_baseline"]["value"]), {k:v["ms_per_launch"] for k,v in d["kernels"].items()})
PY
