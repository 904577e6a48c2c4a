mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-400
python tools/profile_step.py 8192 2 > gpurun_out/step_final.log 2>&1; grep -o "'ms_poa': [0-9.]*\|'ms_total': [0-9.]*" gpurun_out/step_final.log | tr '\n' ' '; echo
timeout 75 python tools/option_fuzz.py 40 2026 gpu > gpurun_out/option_fuzz_gpu.log 2>&1; echo "fuzz exit $?"
grep -c " same" gpurun_out/option_fuzz_gpu.log; grep "DIFFERENT" gpurun_out/option_fuzz_gpu.log | head -n 5 | cut -c1-200
