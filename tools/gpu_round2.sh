mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'poa_kernel|ksw_pair|ksw_ext|chain_dp|partition_kernel' -c 5 \
    -o gpurun_out/prof_full8k -f python tools/profile_step.py 8192 1 > gpurun_out/prof_full8k.log 2>&1
python bench.py --steps 3 --warmup 3 --reads 16384 --no-cpu-baseline > gpurun_out/bench16k.json 2> gpurun_out/bench16k.err
python bench.py --steps 3 --warmup 3 --reads 32768 --no-cpu-baseline > gpurun_out/bench32k.json 2> gpurun_out/bench32k.err
python tools/overlap_test.py 8192 1 > gpurun_out/overlap.log 2>&1
python tools/overlap_test.py 8192 2 >> gpurun_out/overlap.log 2>&1
python tools/overlap_test.py 16384 2 >> gpurun_out/overlap.log 2>&1
python tools/overlap_test.py 16384 4 >> gpurun_out/overlap.log 2>&1
TH_GPU_SHARE=0.5 python tools/overlap_test.py 16384 2 >> gpurun_out/overlap.log 2>&1
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> gpurun_out/overlap.log
