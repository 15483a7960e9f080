#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the three C++ suites: bash tools/gpu_sanitize.sh <tag>
TAG=${1:-r02san}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import sys; sys.path.insert(0,'tests'); from test_cpp_host import build_cpp; build_cpp()"
: > $OUT/sanitizer.txt
for exe in tests/cpp/test_reference_kats tests/cpp/test_container_routes "tests/cpp/test_shard_group 3"; do
  for tool in memcheck racecheck; do
    echo "== $tool $exe" | tee -a $OUT/sanitizer.txt
    CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 400 compute-sanitizer --tool $tool --error-exitcode 7 $exe > $OUT/sanitizer_${tool}_$(basename ${exe%% *}).txt 2>&1
    rc=$?
    { grep -E "checks,|SUMMARY|COMPUTE-SANITIZER" $OUT/sanitizer_${tool}_$(basename ${exe%% *}).txt; echo "exit $rc"; } | tee -a $OUT/sanitizer.txt
  done
done
