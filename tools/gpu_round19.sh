mkdir -p gpurun_out
timeout 900 python tools/parity_at_scale.py > gpurun_out/parity_at_scale.log 2>&1; echo "parity exit $?"
tail -n 3 gpurun_out/parity_at_scale.log | cut -c1-400
nvidia-smi --query-gpu=memory.used --format=csv,noheader
