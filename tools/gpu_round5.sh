mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r5.json 2> gpurun_out/bench_r5.err; echo "bench exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'poa_kernel|chain_dp' -c 2 \
    -o gpurun_out/prof_poa4k -f python tools/profile_step.py 4096 1 > gpurun_out/prof_poa4k.log 2>&1
# phase counters need a -DPOA_PROFILE build; the product build is restored afterwards
TH_NVCC_FLAGS=-DPOA_PROFILE TH_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_prof.log 2>&1
python tools/profile_step.py 8192 2 > gpurun_out/phases_8k.log 2>&1
python tools/profile_step.py 4096 1 short > gpurun_out/phases_short.log 2>&1
python tools/profile_step.py 2048 1 long > gpurun_out/phases_long.log 2>&1
tail -2 gpurun_out/phases_8k.log gpurun_out/phases_short.log gpurun_out/phases_long.log
