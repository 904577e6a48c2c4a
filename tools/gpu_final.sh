mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
timeout 1200 python tools/million_parity.py --check > gpurun_out/million_parity.log 2>&1; echo "million parity exit $?"
tail -n 1 gpurun_out/million_parity.log | cut -c1-700
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'seed_kernel|chain_dp_kernel|rank_kernel|chain_select|partition_kernel|poa_kernel|ksw_pair|ksw_single|ksw_ext|pack_kernel' -c 12 \
    -o gpurun_out/prof_full8k_final -f python tools/profile_step.py 8192 1 > gpurun_out/prof_full8k_final.log 2>&1
tail -n 4 gpurun_out/prof_full8k_final.log | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_final.json"))
print(round(d["value"]), round(d["e2e"]["value"]), d["parity"], d["cpu_baseline"]["value"], {k:v["ms_per_launch"] for k,v in d["kernels"].items()})
PY
