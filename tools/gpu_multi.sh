# usage: bash tools/gpu_multi.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader; nproc
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/sharded_parity.py 6144 > gpurun_out/sharded_parity_n$N.json 2> gpurun_out/sharded_parity_n$N.err; echo "sharded parity exit $?"
tail -n 1 gpurun_out/sharded_parity_n$N.json | cut -c1-300
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
timeout 600 $TR --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref exit $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],1), d["clocks"])
r=json.loads(open("gpurun_out/bench_ref_n$N.json").read().strip().splitlines()[-1]); print("ref", round(r["value"]), r["cpu_baseline"]["cores"])
PY
