#!/bin/bash
# One GPU-box visit, parameterised (replaces the per-round lease scripts).  usage: tools/gpu_job.sh <job> [args...]
#   tests [pytest args]      GPU parity tests
#   step N [steps] [shape]   one device-resident step (tools/profile_step.py)
#   phases N [shape]         rebuild with -DPOA_PROFILE on the box and print the POA phase split
#   ncu NAME N regex [count] ncu --set full capture of kernels matching regex -> gpurun_out/NAME.ncu-rep + summary
#   launches NAME N          ncu launch list (gpu__time_duration) of one step -> gpurun_out/NAME.csv
#   variants N [steps]       time one step with each build_variants/*.so in turn
#   bench [args]             bench.py
#   sanitize TOOL [pytest args]  GPU parity tests under compute-sanitizer (memcheck | racecheck | synccheck | initcheck) -> gpurun_out/r2_sanitizer_TOOL.log
set -u
mkdir -p gpurun_out
job=$1; shift
case $job in
  tests) timeout 1500 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -15 ;;
  step) timeout 600 python tools/profile_step.py "$@" 2>&1 | tail -6 ;;
  phases)
    cp tidehunter_b200/libth_gpu.so /tmp/libth_gpu.so.keep
    TH_NVCC_FLAGS="-DPOA_PROFILE" python -c "import tidehunter_b200.build as b; b.build(force=True)" && timeout 600 python tools/profile_step.py "$@" 2>&1 | tail -4
    cp /tmp/libth_gpu.so.keep tidehunter_b200/libth_gpu.so ;;
  ncu)
    name=$1; n=$2; regex=$3; cnt=${4:-1}; shape=${5:-r2c2}
    timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$regex" -c "$cnt" -o gpurun_out/$name -f python tools/profile_step.py "$n" 1 "$shape" > gpurun_out/$name.log 2>&1
    tail -3 gpurun_out/$name.log | cut -c1-400
    python tools/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/$name.summary.txt 2>&1; cat gpurun_out/$name.summary.txt ;;
  launches)
    name=$1; n=$2
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$name.csv python tools/profile_step.py "$n" 2 > gpurun_out/$name.log 2>&1
    tail -2 gpurun_out/$name.log | cut -c1-300 ;;
  variants) # build_variants/*.so (built here with different flags) take turns as libth_gpu.so: N [steps] [shape]
    cp tidehunter_b200/libth_gpu.so /tmp/libth_gpu.so.keep
    for f in build_variants/*.so; do cp $f tidehunter_b200/libth_gpu.so; echo "== $f"; timeout 600 python tools/profile_step.py "$@" 2>&1 | grep -E "per step"; done
    cp /tmp/libth_gpu.so.keep tidehunter_b200/libth_gpu.so ;;
  bench) timeout 1500 python bench.py "$@" ;;
  sanitize)
    tool=$1; shift
    timeout 1700 compute-sanitizer --tool "$tool" --target-processes all --print-limit 30 --log-file gpurun_out/r2_sanitizer_$tool.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q "$@" 2>&1 | tail -4
    echo "== $tool: $(grep -c 'ERROR SUMMARY' gpurun_out/r2_sanitizer_$tool.log) process(es)"; grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/r2_sanitizer_$tool.log | sort | uniq -c ;;
  *) echo "unknown job $job"; exit 2 ;;
esac
