#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list and full captures of the top kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python tools/profile_step.py 2048 2 > gpurun_out/launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'poa_kernel|ksw_pair|ksw_ext|chain_dp_kernel|seed_kernel|chain_select|rank_kernel' -c 7 \
    -o gpurun_out/prof_full -f python tools/profile_step.py 2048 1 > gpurun_out/prof_full.log 2>&1
ls -la gpurun_out
