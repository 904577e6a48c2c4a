mkdir -p gpurun_out
TH_HOST_TIMING=1 timeout 600 python tools/cli_bench.py 524288 4096 > gpurun_out/cli_bench1.log 2>&1; tail -n 1 gpurun_out/cli_bench1.log | cut -c1-3000
TH_HOST_TIMING=1 TH_CLI_DEVICES=0,1 timeout 600 python tools/cli_bench.py 524288 4096 > gpurun_out/cli_bench2.log 2>&1; tail -n 1 gpurun_out/cli_bench2.log | cut -c1-3000
