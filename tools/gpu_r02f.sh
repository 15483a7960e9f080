#!/bin/bash
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_narrow_division.py tests/test_gpu_scalar_division.py tests/test_gpu_geometry_knobs.py tests/test_gpu_parity.py tests/test_gpu_property.py tests/test_gpu_golden.py -m gpu -q --timeout 600 2>&1 | tail -15 | tee $OUT/pytest_div.txt
echo "== matrix rows"
timeout 600 python tools/dtype_matrix.py --only "div,pow" --dtypes int8,uint8,int16,uint16,float64 --out $OUT/div_matrix.md > $OUT/div_matrix.log 2>&1; tail -3 $OUT/div_matrix.log; grep "^| ew" $OUT/div_matrix.md
