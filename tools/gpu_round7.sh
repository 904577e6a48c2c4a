mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'poa_kernel' -c 1 \
    -o gpurun_out/prof_poa4k_b -f python tools/profile_step.py 4096 1 > gpurun_out/prof_poa4k_b.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mb5.json 2> gpurun_out/bench_mb5.err; echo "bench exit $?"
TH_NVCC_FLAGS=-DPOA_MIN_BLOCKS=8 TH_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_mb8.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mb8.json 2> gpurun_out/bench_mb8.err; echo "bench exit $?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --lanes 2 --chunk 8192 > gpurun_out/bench_mb8_l2.json 2> gpurun_out/bench_mb8_l2.err; echo "bench exit $?"
python - <<'PY'
import json
for f in ("bench_mb5","bench_mb8","bench_mb8_l2"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, round(d["value"]), round(d["e2e"]["value"]), {k:v["ms_per_launch"] for k,v in d["kernels"].items()})
    except Exception as e: print(f, "failed", e)
PY
