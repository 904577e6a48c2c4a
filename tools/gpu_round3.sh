mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l3.json 2> gpurun_out/bench_l3.err; echo "bench exit $?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --reads 32768 --lanes 4 --chunk 4096 > gpurun_out/bench_l4.json 2> gpurun_out/bench_l4.err; echo "bench exit $?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --reads 32768 --lanes 2 --chunk 8192 > gpurun_out/bench_l2.json 2> gpurun_out/bench_l2.err; echo "bench exit $?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --reads 24576 --lanes 3 --chunk 2048 > gpurun_out/bench_l3c2k.json 2> gpurun_out/bench_l3c2k.err; echo "bench exit $?"
