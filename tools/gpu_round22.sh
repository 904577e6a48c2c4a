mkdir -p gpurun_out
timeout 900 python tools/parity_at_scale.py > gpurun_out/parity_at_scale.log 2>&1; echo "parity exit $?"
tail -n 1 gpurun_out/parity_at_scale.log
timeout 900 python tools/config_bench.py > gpurun_out/config_bench.log 2>&1; echo "config bench exit $?"
python - <<'PY'
import json
p=json.load(open("gpurun_out/parity_at_scale.json"))
print([(c["case"], c["identical"], c["bytes"], c["reference_s"], c["ours_s_incl_first_call"]) for c in p["cases"]])
for r in json.load(open("gpurun_out/config_bench.json"))["rows"]:
    print(r["config"], r["e2e_reads_per_s"], r["records_bytes"], r["reference"]["reads_per_s"], r["reference"]["identical_on_sample"], r["speedup_vs_reference"])
PY
