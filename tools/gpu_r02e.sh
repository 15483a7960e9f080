#!/bin/bash
# r02e: 8-bit Power table, f64 Power fast path, signed-lane trims — parity + matrix rows; then full GPU suite.
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== parity: narrow division / power, all-op parity, knobs, property"; timeout 900 python -m pytest tests/test_gpu_narrow_division.py tests/test_gpu_geometry_knobs.py tests/test_gpu_parity.py tests/test_gpu_property.py -m gpu -q --timeout 600 2>&1 | tail -15 | tee $OUT/pytest_div.txt
echo "== (kernel, dtype) matrix: div / rem / pow rows"
timeout 600 python tools/dtype_matrix.py --only "div,rem,pow" --out $OUT/div_matrix.md > $OUT/div_matrix.log 2>&1; tail -3 $OUT/div_matrix.log; cat $OUT/div_matrix.md | tail -80
