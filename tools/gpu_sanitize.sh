#!/bin/bash
# compute-sanitizer over tests/cpp/test_reference_kats (no Python in the way): the reference's transcribed leaf
# tests + the device-path suite (shifted bit windows, ragged consolidate, scalar division, packed 8-bit paths) under
# memcheck, racecheck, initcheck and synccheck.  (memcheck over the pytest suites was tried in r01j: > 16 min, dropped.)
TAG=${1:-r01k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import sys; sys.path.insert(0,'tests'); from test_cpp_host import build_cpp; build_cpp()"
for tool in memcheck racecheck initcheck synccheck; do
  echo "== compute-sanitizer --tool $tool tests/cpp/test_reference_kats"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 tests/cpp/test_reference_kats > $OUT/sanitizer_$tool.txt 2>&1
  echo "exit $?" >> $OUT/sanitizer_$tool.txt; tail -4 $OUT/sanitizer_$tool.txt
done
