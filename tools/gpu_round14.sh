mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'poa_kernel' -c 1 -o gpurun_out/prof_poa8k_final -f python tools/profile_step.py 8192 1 > gpurun_out/prof_poa8k_final.log 2>&1
tail -n 4 gpurun_out/prof_poa8k_final.log | cut -c1-400
python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_final2.json"))
print(round(d["value"]), round(d["e2e"]["value"]), d["parity"]["identical"], d["cpu_baseline"]["value"], {k:v["ms_per_launch"] for k,v in d["kernels"].items()})
PY
