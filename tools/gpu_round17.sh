mkdir -p gpurun_out
for v in "-DPOA_MIN_BLOCKS=8" "-DPART_MIN_BLOCKS=6" "-DPART_MIN_BLOCKS=8" "-DKSW_EXT_MIN_BLOCKS=6" "-DKSW_EXT_MIN_BLOCKS=8" "-DCHAIN_BLOCKS_PER_SM=12" "-DCHAIN_BLOCKS_PER_SM=16"; do
  TH_NVCC_FLAGS="$v" TH_FORCE_BUILD=1 python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_v.log 2>&1
  python tools/profile_step.py 8192 3 > gpurun_out/step_v.log 2>&1; echo "$v"; grep -o "'ms_chain': [0-9.]*\|'ms_partition': [0-9.]*\|'ms_ksw': [0-9.]*\|'ms_total': [0-9.]*" gpurun_out/step_v.log | tr '\n' ' '; echo
done
