#!/bin/bash
# One gpurun call: GPU parity tests, smoke, whole-domain 16-bit division sweep, bench (both arms), division / remainder
# rows of the (kernel, dtype) matrix, ncu launch list + full capture of the top kernels.
# Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh [tag] [full]
#   `full` adds the captures of the batched kernels and of the shifted-window / scalar-division kernels.
TAG=${1:-r01}
FULL=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
{ nvidia-smi; nproc; lscpu | head -20; free -g; } > $OUT/box.txt 2>&1
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
[ -n "$SKIP_SWEEP" ] || { echo "== whole-domain 16-bit division sweep"; timeout 300 python tests/sweep_div16.py 2>&1 | tail -4 | tee $OUT/exhaustive_div16.txt; }
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -3 | tee $OUT/bench_reference.json
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -5 | tee $OUT/bench.json
echo "== (kernel, dtype) matrix: division / remainder rows"
timeout 300 python tools/dtype_matrix.py --only "div,rem" --out $OUT/div_matrix.md > $OUT/div_matrix.log 2>&1; tail -3 $OUT/div_matrix.log
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"reduce_stats|ew_binary|ew_fma|bits_|clear_trailing" -c 400 \
    --csv --log-file $OUT/launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-supertable > $OUT/ncu_launches.log 2>&1
tail -2 $OUT/ncu_launches.log
echo "== ncu full: reduce"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:reduce_stats -s 3 -c 2 -f -o $OUT/prof_reduce \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-secondary > $OUT/ncu_reduce.log 2>&1
tail -2 $OUT/ncu_reduce.log
echo "== ncu full: ew f64 masked add"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ew_binary -s 3 -c 2 -f -o $OUT/prof_ew \
    python bench.py --steps 3 --warmup 3 --rows 67108864 --no-e2e --no-cpu --no-supertable > $OUT/ncu_ew.log 2>&1
tail -2 $OUT/ncu_ew.log
if [ "$FULL" = "full" ]; then
echo "== ncu full: batched kernels (C5 SuperTable)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:batch_kernel -s 8 -c 4 -f -o $OUT/prof_batch \
    python bench.py --steps 3 --warmup 3 --rows 67108864 --no-e2e --no-cpu > $OUT/ncu_batch.log 2>&1
tail -2 $OUT/ncu_batch.log
echo "== ncu full: shifted bit windows + i64 / scalar (multiplicative inverse)"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"bits_shift_kernel|ew_binary_kernel<long, long, long, (mnr::)?V32, 4," -s 4 -c 3 -f -o $OUT/prof_shift_sdiv \
    python bench.py --steps 3 --warmup 3 --rows 67108864 --no-e2e --no-cpu --no-supertable > $OUT/ncu_shift_sdiv.log 2>&1
tail -2 $OUT/ncu_shift_sdiv.log
fi
ls -la $OUT
