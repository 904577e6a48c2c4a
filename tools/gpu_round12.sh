mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
python tools/profile_step.py 8192 2 > gpurun_out/step_a.log 2>&1; grep -o "'ms_poa': [0-9.]*" gpurun_out/step_a.log
python tools/profile_step.py 4096 1 short > gpurun_out/step_short.log 2>&1; grep -o "'ms_poa': [0-9.]*" gpurun_out/step_short.log
python tools/profile_step.py 2048 1 long > gpurun_out/step_long.log 2>&1; grep -o "'ms_poa': [0-9.]*" gpurun_out/step_long.log
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r12.json 2> gpurun_out/bench_r12.err; echo "bench exit $?"
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:poa_kernel -c 1 --csv --log-file gpurun_out/poa_inst.csv python tools/profile_step.py 4096 1 > /dev/null 2>&1
tail -n 1 gpurun_out/poa_inst.csv | awk -F'","' '{print $(NF-2), $NF}'
TH_NVCC_FLAGS=-DPOA_PROFILE TH_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_prof.log 2>&1
python tools/profile_step.py 8192 2 > gpurun_out/phases_8k.log 2>&1
tail -n 2 gpurun_out/phases_8k.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r12.json"))
print(round(d["value"]), round(d["e2e"]["value"]), {k:v["ms_per_launch"] for k,v in d["kernels"].items()})
PY
