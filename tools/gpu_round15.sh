mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python tools/cli_bench.py 262144 8192 > gpurun_out/cli_bench1.log 2>&1; tail -n 1 gpurun_out/cli_bench1.log | cut -c1-700
TH_CLI_DEVICES=0,1 timeout 600 python tools/cli_bench.py 262144 8192 > gpurun_out/cli_bench2.log 2>&1; tail -n 1 gpurun_out/cli_bench2.log | cut -c1-700
