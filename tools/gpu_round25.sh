mkdir -p gpurun_out
for r in 65536 49152; do
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --reads $r > gpurun_out/bench_reads$r.json 2> gpurun_out/bench_reads$r.err; echo "bench $r exit $?"
done
python - <<'PY'
import json
for r in (65536, 49152):
    d=json.load(open("gpurun_out/bench_reads%d.json"%r))
    print(r, round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], {k:v["ms_per_launch"] for k,v in d["kernels"].items()})
PY
