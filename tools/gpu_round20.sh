mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 900 python tools/parity_at_scale.py > gpurun_out/parity_at_scale.log 2>&1; echo "parity exit $?"
tail -n 2 gpurun_out/parity_at_scale.log | cut -c1-300
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r20.json 2> gpurun_out/bench_r20.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r20.json"))
print(round(d["value"]), round(d["e2e"]["value"]))
p=json.load(open("gpurun_out/parity_at_scale.json"))
print([(c["case"], c["identical"], c["reference_s"], c["ours_s_incl_first_call"]) for c in p["cases"]])
PY
