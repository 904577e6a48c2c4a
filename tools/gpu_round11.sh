mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
python tools/profile_step.py 8192 2 > gpurun_out/step_a.log 2>&1; grep -o "'ms_poa': [0-9.]*" gpurun_out/step_a.log
python tools/profile_step.py 16384 2 > gpurun_out/step_b.log 2>&1; grep -o "'ms_poa': [0-9.]*" gpurun_out/step_b.log
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r11.json 2> gpurun_out/bench_r11.err; echo "bench exit $?"
TH_HOST_TIMING=1 python tools/e2e_probe.py 32768 4096 4 3 > gpurun_out/e2e_probe.log 2>&1
tail -n 8 gpurun_out/e2e_probe.log
TH_NVCC_FLAGS=-DPOA_PROFILE TH_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_prof.log 2>&1
python tools/profile_step.py 8192 2 > gpurun_out/phases_8k.log 2>&1
tail -n 2 gpurun_out/phases_8k.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r11.json"))
print(round(d["value"]), round(d["e2e"]["value"]), {k:v["ms_per_launch"] for k,v in d["kernels"].items()})
PY
