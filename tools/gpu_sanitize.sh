#!/bin/bash
# compute-sanitizer over the C++ test binaries (no Python in the way): the reference's transcribed leaf
# tests + the device-path suite (shifted bit windows, ragged consolidate, scalar division, packed 8-bit paths, 8/16-bit
# division through the f32 pipe, float remainder) and the router / container routes (batched launches), under memcheck,
# racecheck, initcheck and synccheck.  (memcheck over the pytest suites was tried in r01j: > 16 min, dropped.)
TAG=${1:-r01k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import sys; sys.path.insert(0,'tests'); from test_cpp_host import build_cpp; build_cpp()"
: > $OUT/sanitizer.txt
for exe in tests/cpp/test_reference_kats tests/cpp/test_container_routes; do
  for tool in memcheck racecheck initcheck synccheck; do
    echo "== $tool $exe" | tee -a $OUT/sanitizer.txt
    timeout 300 compute-sanitizer --tool $tool --error-exitcode 7 $exe > $OUT/sanitizer_${tool}_$(basename $exe).txt 2>&1
    rc=$?
    { grep -E "checks,|SUMMARY|COMPUTE-SANITIZER" $OUT/sanitizer_${tool}_$(basename $exe).txt; echo "exit $rc"; } | tee -a $OUT/sanitizer.txt
  done
done
