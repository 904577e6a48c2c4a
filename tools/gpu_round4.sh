mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python tools/parity_at_scale.py > gpurun_out/parity_at_scale.log 2>&1; echo "parity exit $?"
tail -2 gpurun_out/parity_at_scale.log
timeout 600 python bench.py > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r1b.json"))
print({k:d.get(k) for k in ("value","ms_per_step","parity","cpu_baseline")}, d["e2e"]["value"], {k:v["ms_per_launch"] for k,v in d["kernels"].items()})
PY
