mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r9.json 2> gpurun_out/bench_r9.err; echo "bench exit $?"
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --reads 65536 > gpurun_out/bench_r9_64k.json 2> gpurun_out/bench_r9_64k.err; echo "bench exit $?"
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --reads 49152 --lanes 3 > gpurun_out/bench_r9_l3.json 2> gpurun_out/bench_r9_l3.err; echo "bench exit $?"
python - <<'PY'
import json
for f in ("bench_r9","bench_r9_64k","bench_r9_l3"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, round(d["value"]), round(d["e2e"]["value"]), {k:v["ms_per_launch"] for k,v in d["kernels"].items()}, {k:v["ms_per_launch"] for k,v in d["kernels_overlapped"].items()})
    except Exception as e: print(f, "failed", e)
PY
nvidia-smi --query-gpu=memory.used --format=csv,noheader
