#!/bin/bash
# compute-sanitizer over the small-shape suites: memcheck, racecheck, initcheck, synccheck on the C++ KAT
# binary (no Python in the way), memcheck on the golden-vector pytest, plus the batch-kernel ncu captures.
TAG=${1:-r01k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c "import sys; sys.path.insert(0,'tests'); from test_cpp_host import build_cpp; build_cpp()"
for tool in memcheck racecheck initcheck synccheck; do
  echo "== compute-sanitizer --tool $tool tests/cpp/test_reference_kats"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 tests/cpp/test_reference_kats > $OUT/sanitizer_$tool.txt 2>&1
  echo "exit $?" >> $OUT/sanitizer_$tool.txt; tail -4 $OUT/sanitizer_$tool.txt
done
echo "== memcheck: golden vectors + eq_mask + concat through pytest"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_golden.py tests/test_gpu_concat.py tests/test_eq_mask.py tests/test_gpu_broadcast_routes.py -m gpu -q -x > $OUT/sanitizer_memcheck_pytest.txt 2>&1
echo "exit $?" >> $OUT/sanitizer_memcheck_pytest.txt; tail -6 $OUT/sanitizer_memcheck_pytest.txt
echo "== ncu: batched reduce (C5)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reduce_stats_batch_kernel -s 21 -c 4 -f -o $OUT/prof_batch_reduce \
    python bench.py --steps 3 --warmup 3 --rows 67108864 --no-e2e --no-cpu > $OUT/ncu_batch_reduce.log 2>&1; tail -1 $OUT/ncu_batch_reduce.log
echo "== ncu: batched element-wise (C5)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ew_binary_batch_kernel -c 4 -f -o $OUT/prof_batch_ew \
    python bench.py --steps 3 --warmup 3 --rows 67108864 --no-e2e --no-cpu > $OUT/ncu_batch_ew.log 2>&1; tail -1 $OUT/ncu_batch_ew.log
