#!/bin/bash
# Short follow-up call: C++ suites natively, sanitizers over them, one full ncu capture of the i64 / scalar kernel.
TAG=${1:-r01zz}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== C++ suites"; timeout 300 python -m pytest tests/test_cpp_host.py -m gpu -q --timeout 200 2>&1 | tail -4 | tee $OUT/pytest_cpp.txt
bash tools/gpu_sanitize.sh $TAG
echo "== ncu full: i64 / scalar through the multiplicative inverse"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"ew_binary_kernel<long, long, long, (mnr::)?V32, 4," -s 3 -c 1 -f -o $OUT/prof_sdiv \
    python bench.py --steps 3 --warmup 3 --rows 67108864 --no-e2e --no-cpu --no-supertable > $OUT/ncu_sdiv.log 2>&1
tail -2 $OUT/ncu_sdiv.log
ls -la $OUT
