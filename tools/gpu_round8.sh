mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
TH_GPU_DEBUG=1 python tools/profile_step.py 8192 2 > gpurun_out/step_u4.log 2>&1; echo "u4 mb8"; grep -o "'ms_poa': [0-9.]*" gpurun_out/step_u4.log; grep -c "retried" gpurun_out/step_u4.log
TH_GPU_DEBUG=1 python tools/profile_step.py 4096 1 short > gpurun_out/step_short.log 2>&1; grep -o "'ms_poa': [0-9.]*" gpurun_out/step_short.log; grep "retried" gpurun_out/step_short.log | head -n 2
TH_GPU_DEBUG=1 python tools/profile_step.py 2048 1 long > gpurun_out/step_long.log 2>&1; grep -o "'ms_poa': [0-9.]*" gpurun_out/step_long.log; grep "retried" gpurun_out/step_long.log | head -n 2
for v in "-DPOA_SETUP_U=2" "-DPOA_SETUP_U=1" "-DPOA_SETUP_U=2 -DPOA_MIN_BLOCKS=7" "-DPOA_SETUP_U=4 -DPOA_MIN_BLOCKS=6"; do
  TH_NVCC_FLAGS="$v" TH_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_v.log 2>&1
  python tools/profile_step.py 8192 2 > gpurun_out/step_v.log 2>&1; echo "$v"; grep -o "'ms_poa': [0-9.]*" gpurun_out/step_v.log
done
TH_NVCC_FLAGS=-DPOA_PROFILE TH_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_prof.log 2>&1
python tools/profile_step.py 8192 2 > gpurun_out/phases_8k.log 2>&1
tail -n 2 gpurun_out/phases_8k.log
