mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python tools/cli_bench.py 131072 8192 > gpurun_out/cli_bench.log 2>&1; tail -n 3 gpurun_out/cli_bench.log | cut -c1-900
