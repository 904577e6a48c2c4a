mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
python tools/profile_step.py 8192 2 > gpurun_out/step_mb5.log 2>&1; grep -o "'ms_poa': [0-9.]*" gpurun_out/step_mb5.log
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:poa_kernel -c 1 --csv --log-file gpurun_out/poa_inst.csv python tools/profile_step.py 4096 1 > /dev/null 2>&1
tail -n 2 gpurun_out/poa_inst.csv
for mb in 6 8; do
  TH_NVCC_FLAGS=-DPOA_MIN_BLOCKS=$mb TH_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_mb$mb.log 2>&1
  python tools/profile_step.py 8192 2 > gpurun_out/step_mb$mb.log 2>&1; echo "mb $mb"; grep -o "'ms_poa': [0-9.]*" gpurun_out/step_mb$mb.log
done
TH_NVCC_FLAGS=-DPOA_PROFILE TH_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_prof.log 2>&1
python tools/profile_step.py 8192 2 > gpurun_out/phases_8k.log 2>&1
tail -n 2 gpurun_out/phases_8k.log
