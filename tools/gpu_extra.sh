#!/bin/bash
# Short follow-up call: full ncu captures of the division kernels (i64 / scalar through the multiplicative inverse,
# i16 column division through the f32 pipe, f64 remainder through one division + fma).
TAG=${1:-r01zz}; OUT=gpurun_out/$TAG; mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
echo "== ncu full: i64 / scalar"
timeout 200 $NCU -k regex:"ew_binary_kernel<long, long, long, mnr::V32, \(int\)4" -s 3 -c 1 -f -o $OUT/prof_sdiv \
    python bench.py --steps 3 --warmup 3 --rows 67108864 --no-e2e --no-cpu --no-supertable > $OUT/ncu_sdiv.log 2>&1
tail -2 $OUT/ncu_sdiv.log
echo "== ncu full: i16 / i16 (narrow_quot)"
timeout 200 $NCU -k regex:"ew_binary_kernel<short, short, short, mnr::V16, \(int\)1" -s 3 -c 1 -f -o $OUT/prof_div16 \
    python tools/dtype_matrix.py --only "ew div two masks" --dtypes int16 > $OUT/ncu_div16.log 2>&1
tail -2 $OUT/ncu_div16.log
echo "== ncu full: f64 % f64 (fast_fmod)"
timeout 200 $NCU -k regex:"ew_binary_kernel<double, double, double, mnr::V16, \(int\)3" -s 3 -c 1 -f -o $OUT/prof_rem64 \
    python tools/dtype_matrix.py --only "ew rem two masks" --dtypes float64 > $OUT/ncu_rem64.log 2>&1
tail -2 $OUT/ncu_rem64.log
ls -la $OUT
