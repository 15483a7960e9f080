#!/bin/bash
# Round-2 confirmation on the final build (one 1-GPU call): full GPU suite, smoke, whole-domain 16-bit division sweep, both
# bench arms, the whole (kernel, dtype) matrix, ncu launch list + full captures of the top kernels, compute-sanitizer over
# the C++ suites (group test with virtual ranks included).
# Usage: bash tools/gpu_r02_final.sh [tag] [nosan]
TAG=${1:-r02z}; NOSAN=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
{ nvidia-smi; nproc; lscpu | head -20; free -g; } > $OUT/box.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== whole-domain 16-bit division sweep"; timeout 300 python tests/sweep_div16.py 2>&1 | tail -4 | tee $OUT/exhaustive_div16.txt
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_reference.json | cut -c1-300
echo "== bench"; timeout 600 python bench.py > $OUT/bench.log 2>&1; tail -1 $OUT/bench.log > $OUT/bench.json; cut -c1-400 $OUT/bench.json
echo "== bench, driver-like (steps 20)"; timeout 300 python bench.py --steps 20 --warmup 3 --no-secondary --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_steps20.json | cut -c1-300
echo "== (kernel, dtype) matrix"
timeout 600 python tools/dtype_matrix.py --out $OUT/dtype_matrix.md > $OUT/dtype_matrix.log 2>&1; tail -2 $OUT/dtype_matrix.log
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"reduce_stats|ew_binary|ew_fma|bits_|clear_trailing|fold_exchange" -c 400 \
    --csv --log-file $OUT/launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --rows 268435456 > $OUT/ncu_launches.log 2>&1
tail -2 $OUT/ncu_launches.log
echo "== ncu full: reduce (the bench's kernel: exchange form, world 1)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:reduce_stats_kernel -s 30 -c 2 -f -o $OUT/prof_reduce \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-secondary > $OUT/ncu_reduce.log 2>&1
tail -2 $OUT/ncu_reduce.log
ncu -i $OUT/prof_reduce.ncu-rep --page raw --csv > $OUT/prof_reduce_raw.csv 2>/dev/null; rm -f $OUT/prof_reduce.ncu-rep
echo "== ncu full: ew f64 masked add / batched kernels (configs[2], configs[4])"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ew_binary_kernel|batch_kernel" -s 6 -c 6 -f -o $OUT/prof_ew \
    python bench.py --steps 3 --warmup 3 --rows 67108864 --no-e2e --no-cpu > $OUT/ncu_ew.log 2>&1
tail -2 $OUT/ncu_ew.log
ncu -i $OUT/prof_ew.ncu-rep --page raw --csv > $OUT/prof_ew_raw.csv 2>/dev/null; rm -f $OUT/prof_ew.ncu-rep
echo "== ncu full: packed u8 division + 8-bit power table + f64 power"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ew_binary_kernel -c 12 -f -o $OUT/prof_narrow \
    python tools/dtype_matrix.py --only "div two,pow" --dtypes uint8,int8,float64 --gib 0.25 --iters 1 --out $OUT/ncu_narrow_matrix.md > $OUT/ncu_narrow.log 2>&1
tail -2 $OUT/ncu_narrow.log
ncu -i $OUT/prof_narrow.ncu-rep --page raw --csv > $OUT/prof_narrow_raw.csv 2>/dev/null; rm -f $OUT/prof_narrow.ncu-rep
if [ -z "$NOSAN" ]; then
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
fi
du -sh $OUT; ls -la $OUT | tail -40
